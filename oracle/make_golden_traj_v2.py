"""Golden trajectory of the reference's cycle-per-step path (`SBR-v2`), authoring container only.

What `SBR_model_FB.run` returns as `t`, `x` (every phase's output points stacked, SBR_model_FB.py:71-86 and the like) and as
its kla3 / kla5 / kla8 arrays, captured from the UNMODIFIED reference during one `SbrEnv2.step` (oracle/make_golden.py's
spy on `_next_observation`), default LSODA and LSODA at rtol = atol = 1e-12.  Stored: the state at the END of every PID interval
of phases 1-5 and 8 (the points sbr_cycle_v2_traj records), the post-draw state, and the three KLa arrays.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
import ref_shim  # noqa: E402
import sbr_oracle as O  # noqa: E402


def interval_ends(t, x):
    """Indices of the interval-end columns of `x` [14, P] and of the first draw-phase column.  Phase k in (1..5, 8) holds
    1 + n_int (pts - 1) columns (its start, then pts - 1 per PID interval); the settle and draw phases hold
    int((t1 - t0) / t_delta) columns each."""
    sched = O.cycle_schedule()
    idx, col, post_draw = [], 0, None
    for k, (t0, t1) in enumerate(sched):
        if k in (5, 6):
            npts = int((t1 - t0) / O.DT)
            if k == 6:
                post_draw = col
            col += npts
            continue
        grid, pts = O.phase_grid(t0, t1)
        for i in range(len(grid) - 1):
            idx.append(col + (i + 1) * (pts[i] - 1))
        col += 1 + (len(grid) - 1) * (pts[0] - 1)
    assert col == x.shape[1] == len(t), (col, x.shape, len(t))
    return np.array(idx), post_draw


def main():
    ref_shim.load_reference()
    out = {"versions": MG.VERSIONS}
    for tag, tight in (("default", False), ("tight", True)):
        case = None
        import gym_SBR.envs.gym_SBR_env2 as m
        np.random.seed(0)
        env = m.SbrEnv2()
        cap = {}
        orig = env._next_observation

        def spy(*a, **k):
            r = orig(*a, **k)
            cap["run"] = r
            return r

        env._next_observation = spy
        action = np.array([0.25, 0.5, 0.75])
        with ref_shim.quiet():
            env.reset()
            influent = np.array(m.influent_mixed, dtype=float)
            if tight:
                with MG.tight_odeint():
                    env.step(action)
            else:
                env.step(action)
        t, x, x_last = cap["run"][0], np.array(cap["run"][1], dtype=float), cap["run"][2]
        kla3, kla5, kla8 = cap["run"][15], cap["run"][16], cap["run"][17]
        idx, post = interval_ends(np.array(t, dtype=float), x)
        out.update({tag + "_t": np.array(t, dtype=float)[idx], tag + "_x": x[:, idx].T.copy(),
                    tag + "_x_post_draw": x[:, post].copy(), tag + "_x_last": np.array(x_last, dtype=float),
                    tag + "_kla3": np.array(kla3, dtype=float), tag + "_kla5": np.array(kla5, dtype=float),
                    tag + "_kla8": np.array(kla8, dtype=float)})
        out["influent"], out["action"] = influent, action
        print(tag, "points", x.shape[1], "interval ends", len(idx), "x_last[8:11]", np.array(x_last)[8:11], flush=True)
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "sbr_v2_traj_seed0.npz"), **out)


if __name__ == "__main__":
    main()
