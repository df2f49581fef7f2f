"""Golden episodes of the reference `SBR-v4` env (gym_SBR_env4.py), authoring container only.

DISCLOSURE: `SbrEnv4.step` as shipped raises TypeError on numpy >= 1.18, because it passes a FLOAT `num` to
np.linspace in four places (gym_SBR_env4.py:286,921,982,1207).  The file dates from the numpy < 1.18 era, where linspace
truncated a float `num` with int() (and a DeprecationWarning).  This harness runs the UNMODIFIED reference source with
that historical behaviour restored FOR THAT MODULE ONLY: the module-level name `np` of gym_SBR_env4 is replaced by a
proxy whose `linspace` casts `num` with int() and forwards everything else to numpy.  No reference source is copied or
edited.  Any parity claim for SBR-v4 is against this shimmed run.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


class _NumpyPre118(object):
    """numpy with linspace(num=float) -> int(num), as numpy < 1.18 behaved."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def linspace(start, stop, num=50, *a, **k):
        return np.linspace(start, stop, int(num), *a, **k)


def load_env4():
    ref_shim.load_reference()
    import gym_SBR.envs.gym_SBR_env4 as m
    m.np = _NumpyPre118()
    return m


def action_plan(kind, rng, n=700):
    if kind == "zero":
        return np.zeros((n, 1))
    if kind == "up":                       # ramp the DO set-point up early, then hold
        a = np.zeros((n, 1)); a[:40] = 0.05
        return a
    if kind == "random":
        return rng.uniform(-1, 1, (n, 1))
    if kind == "walk":
        return np.clip(0.2 * rng.randn(n, 1) + 0.02, -1, 1)
    raise ValueError(kind)


def run_episode(seed, kind):
    m = load_env4()
    rng = np.random.RandomState(2000 + seed)
    plan = action_plan(kind, rng)
    np.random.seed(seed)
    env = m.SbrEnv4()
    rec = dict(state=[], reward=[], done=[], action=[], t=[], batch_type=[], u=[], kla=[])
    with ref_shim.quiet(), warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        obs0 = env.reset()
        influent = np.array(m.influent_mixed, dtype=float)
        k = 0
        while True:
            state, reward, done, info = env.step(float(plan[k, 0]))
            rec["state"].append(np.array(state, dtype=float).reshape(-1))
            rec["reward"].append(float(reward)); rec["done"].append(bool(done)); rec["action"].append(float(plan[k, 0]))
            rec["t"].append(float(m.t)); rec["batch_type"].append(int(m.batch_type)); rec["u"].append(float(m.u))
            rec["kla"].append(float(m.Kla[-1]))
            k += 1
            if done or k >= 700:
                break
        nwarn = len([w for w in wlist if "ODEint" in str(w.category)])
    out = {k_: np.array(v) for k_, v in rec.items()}
    out.update(seed=seed, kind=kind, reset_obs=np.array(obs0, dtype=float).reshape(-1), influent=influent,
               n_steps=k, Qw=float(m.Qw), eff=np.array(m.eff_component, dtype=float), kla_sum=float(sum(m.Kla)),
               n_kla=len(m.Kla), odeint_warnings=nwarn)
    return out


def make_v4(out_dir, versions):
    for seed, kind in [(0, "zero"), (1, "up"), (2, "walk"), (3, "random"), (4, "walk"), (5, "up")]:
        ep = run_episode(seed, kind)
        print("v4 seed %d %-6s steps %d sumR %.9g Qw %.6g finite %s warn %d" % (
            seed, kind, ep["n_steps"], ep["reward"].sum(), ep["Qw"], bool(np.isfinite(ep["state"]).all()),
            ep["odeint_warnings"]), flush=True)
        np.savez_compressed(os.path.join(out_dir, "sbr_v4_seed%d_%s.npz" % (seed, kind)), versions=versions, **ep)


if __name__ == "__main__":
    import scipy
    out = os.path.join(os.path.dirname(HERE), "tests", "golden")
    make_v4(out, np.array([np.__version__, scipy.__version__, sys.version.split()[0]]))
