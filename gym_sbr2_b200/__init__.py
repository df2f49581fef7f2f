"""gym_sbr2_b200 -- B200-native batched sequencing-batch-reactor simulator behind the gym-SBR API.

    import gym_sbr2_b200 as sbr
    env = sbr.make("SBR-v2")                    # per-instance Gym env (batch of one on cuda:0)
    vec = sbr.SbrV2VecEnv(1 << 20, "cuda:0")    # the vectorised drop-in: torch CUDA tensors in and out
    (likewise SbrOsVecEnv, SbrV4VecEnv, SbrCntVecEnv(kind, n) for SBROS-v1, SBR-v4, SBRCnt-v0/1/2 / SBRCntMA-v1 / SBROS-v2,
     SbrIlcVecEnv for the batch-to-batch feed-forward KLa path of SBR-v0, SbrV1VecEnv for SBR-v1)

Importing the package registers the reference's ten env ids (with gym / gymnasium too when installed).
"""
__version__ = "0.3.0"

from .registration import ENV_TABLE, UnsupportedEnvError, make, register, registry, spec_ids  # noqa: F401


def __getattr__(name):
    # the vector envs import torch; keep `import gym_sbr2_b200` light
    if name in ("SbrV2VecEnv", "SbrOsVecEnv", "SbrV4VecEnv"):
        from . import vec_env
        return getattr(vec_env, name)
    if name == "SbrCntVecEnv":
        from . import cnt
        return cnt.SbrCntVecEnv
    if name in ("SbrIlcVecEnv", "SbrV1VecEnv"):
        from . import ilc
        return getattr(ilc, name)
    raise AttributeError(name)
