"""Single-environment Gym classes under the reference's class names (gym_SBR/envs/__init__.py:1-10).

`SbrEnv2`, `SbrOS` and `SbrEnv4` are thin wrappers over a batch of ONE env of the CUDA vector envs, returning numpy / Python
values exactly shaped like the reference's; "identical seeds" means `np.random.seed(s)` before `reset()` as in the
reference (global numpy RNG, buffer_tank3.py:68) -- the influent draw consumes the same random numbers.
`SbrCnt0`, `SbrCnt1`, `SbrCnt2`, `SbrCntMA1` and `SbrOS1` wrap the CUDA path of the five ids whose reference `step()` dies
in its reward module (repaired reward, disclosed in oracle/make_golden_cnt.py).  `SbrEnv` (`SBR-v0`) serves the batch-to-batch feed-forward
KLa path: everything its reference `step()` does before the reward call that cannot run (disclosed in
oracle/make_golden_ilc.py).  `SbrEnv1` (`SBR-v1`) is the same plant under the feedback PID alone: its cycle
(SBR_model_FBc_implemented.run) is served and pinned, its reward call dies in the reference like `SBR-v0`'s.
"""
from .single import SbrCnt0, SbrCnt1, SbrCnt2, SbrCntMA1, SbrEnv, SbrEnv1, SbrEnv2, SbrEnv4, SbrOS, SbrOS1

__all__ = ["SbrEnv", "SbrEnv1", "SbrEnv2", "SbrEnv4", "SbrCnt0", "SbrCnt1", "SbrCnt2", "SbrCntMA1", "SbrOS", "SbrOS1"]
