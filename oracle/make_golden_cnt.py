"""Golden episodes of the five interval-per-step ids whose reference `step()` dies in its reward module
(`SBRCnt-v0/1/2`, `SBRCntMA-v1`, `SBROS-v2`), authoring container only.

DISCLOSURE.  Every one of these envs calls `module_reward_continuous1.sbr_reward`, which cannot run as shipped:
  * module_reward_continuous1.py:32-39 test the name `So`, which is never bound (`so = x_out[8]` is, :6)  -> NameError;
  * :61 adds `r_snh`, whose only assignment sits inside a string literal (:43-49)                         -> NameError;
  * on the `done` branch (:22-24) `r_e` is never bound either.
The env modules themselves are run UNMODIFIED; only the name `sbr_reward` they imported is rebound to the repaired
function below, which keeps the reference's own threshold table and binds the three missing names the way the
sibling module module_reward_continuous.py:4-65 does:
  So    := so (= x_out[8], :6)
  r_snh := 0 while the cycle reacts; at `done`, 0 if eff[3] (effluent Snh) < 4 else -246 when the env passes the effluent
           vector (module_reward_continuous.py:40-51), 0 when it passes the scalar 0
  r_e   := 0 on the `done` branch
Rewards of these ids are therefore "parity by construction" (SURVEY.md 8f); states, observations, `done` and every
controller quantity come from the reference's own code.

Several of these envs leave the physical regime on their own: the carbon controllers of SbrCnt2 / SbrCntMA1 / SbrOS1 have
no upper clamp and the error sign `sp - cv` (gym_SBR_continuous2.py:947-965, gym_SBR_continuous_MA1.py:963-984), so whenever
the set-point is above the measured value the dosing flow integrates up to ~1e3 m3/d and the reactor volume to hundreds
of m3.  The fixtures hold such episodes too (`physical` False); parity tests use them up to the last reacting step only.
"""
import importlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

KINDS = {
    "cnt0": ("gym_SBR_continuous0", "SbrCnt0"),
    "cnt1": ("gym_SBR_continuous1", "SbrCnt1"),
    "cnt2": ("gym_SBR_continuous2", "SbrCnt2"),
    "ma1": ("gym_SBR_continuous_MA1", "SbrCntMA1"),
    "os2": ("gym_SBR_oneshot1", "SbrOS1"),
}


def repaired_reward(x_out, u_t, done, eff):
    """module_reward_continuous1.sbr_reward with its three unbound names bound (see the module docstring)."""
    so = x_out[8]
    if done:
        r_e = 0
        r_snh = 0
        if not np.isscalar(eff):
            r_snh = 0 if eff[3] < 4 else -246
    else:
        if so < 1.5:
            r_e = -100
        elif 2.5 < so < 3.5:
            r_e = 0
        elif 3.5 <= so < 5:
            r_e = -10
        elif 5 <= so:
            r_e = -50
        else:
            r_e = 10
        r_snh = 0
    return r_snh + r_e


def load_env(kind):
    ref_shim.load_reference()
    modname, cls = KINDS[kind]
    m = importlib.import_module("gym_SBR.envs." + modname)
    m.sbr_reward = repaired_reward
    return m, getattr(m, cls)


def action_plan(kind, plan, rng, n=520):
    """[n, 2] actions (column 1 is only read by os2)."""
    a = np.zeros((n, 2))
    if kind == "os2":
        if plan == "const":
            a[:, 0], a[:, 1] = 2.0, 0.0
        elif plan == "walk":
            a[:, 0] = np.clip(2.0 + np.cumsum(0.05 * rng.randn(n)), 0.5, 6.0)
            a[:, 1] = 0.0
        elif plan == "lowdose":
            a[:, 0] = rng.uniform(1.0, 3.0, n)
            a[:, 1] = 0.3
        elif plan == "dose":
            a[:, 0], a[:, 1] = 2.0, 5.0
        else:
            raise ValueError(plan)
        return a
    if plan == "zero":
        pass
    elif plan == "up":                      # raise the DO set-point to ~2 g/m3 early, then hold
        if kind == "cnt0":
            a[:40, 0] = 0.05
        elif kind == "ma1":
            a[:2, 0] = -1.0                 # brings the carbon set-point from 2 to 0 before it can dose
            a[60:64, 0] = 0.5               # first aerobic steps
            a[300:304, 0] = 0.0
        elif kind == "cnt2":
            a[0, 0] = -2.0                  # carbon set-point 2 -> 0 (it only moves in the first step)
            a[1:5, 0] = 0.5
        else:
            a[:4, 0] = 0.5
    elif plan == "walk":
        scale = 0.02 if kind == "cnt0" else 0.2
        a[:, 0] = np.clip(scale * rng.randn(n) + 0.1 * scale, -1, 1)
        if kind == "ma1":
            a[:2, 0] = -1.0
        if kind == "cnt2":
            a[0, 0] = -2.0
    elif plan == "random":
        a[:, 0] = rng.uniform(-1, 1, n) * (0.05 if kind == "cnt0" else 1.0)
        if kind == "ma1":
            a[:2, 0] = -1.0
        if kind == "cnt2":
            a[0, 0] = -2.0
    elif plan == "dose":                    # leaves the carbon set-point where reset puts it: the controller runs away
        a[:4, 0] = 0.5
    else:
        raise ValueError(plan)
    return a


def run_episode(kind, seed, plan):
    m, cls = load_env(kind)
    rng = np.random.RandomState(3000 + seed)
    acts = action_plan(kind, plan, rng)
    np.random.seed(seed)
    env = cls()
    keys = ("obs", "state15", "reward", "done", "action", "t", "x_end", "x_cont", "u_do", "u_ec", "kla", "ec")
    rec = {k: [] for k in keys}
    with ref_shim.quiet(), warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        obs0 = env.reset()
        influent = np.array(m.influent_mixed, dtype=float)
        x_fill = np.array(m.x_out[-1], dtype=float)
        k = 0
        while True:
            a = acts[k]
            out = env.step(a.copy() if kind == "os2" else np.array([a[0]]))
            if kind == "os2":
                obs, state, reward, done, _ = out
                rec["obs"].append(np.concatenate([np.asarray(obs[0], float), np.asarray(obs[1], float)]))
                rec["state15"].append(np.asarray(state, float).reshape(-1))
            else:
                obs, reward, done, _ = out
                rec["obs"].append(np.asarray(obs, float).reshape(-1))
                rec["state15"].append(np.zeros(0))
            rec["reward"].append(float(reward)); rec["done"].append(bool(done)); rec["action"].append(a.copy())
            rec["t"].append(float(m.t))
            rec["x_end"].append(np.array(m.x_out[-1], dtype=float))      # last row of the step's x_out
            rec["x_cont"].append(np.array(m.x_t[-1], dtype=float))       # the state the next step continues from
            u_do = m.u if kind in ("cnt0", "cnt1") else m.u_DO
            rec["u_do"].append(float(np.asarray(u_do).reshape(-1)[0]))
            rec["u_ec"].append(float(np.asarray(getattr(m, "u_EC", 0.0)).reshape(-1)[0]))
            rec["kla"].append(float(m.Kla[-1]))
            rec["ec"].append(float(m.EC[-1]) if hasattr(m, "EC") and len(m.EC) else 0.0)
            k += 1
            if done or k >= len(acts):
                break
        nwarn = len([w for w in wlist if "ODEint" in str(w.category)])
    out = {k_: np.array(v) for k_, v in rec.items()}
    vmax = float(np.max(out["x_cont"][:, 0]))
    qw = m.Qw if np.isscalar(m.Qw) or isinstance(m.Qw, (float, np.floating)) else np.nan
    out.update(kind=kind, seed=seed, plan=plan, reset_obs=_flat_obs(obs0, kind), influent=influent, x_fill=x_fill,
               n_steps=k, Qw=float(qw), v_max=vmax, physical=bool(vmax < 1.4 and np.isfinite(out["x_cont"]).all()),
               odeint_warnings=nwarn)
    return out


def _flat_obs(o, kind):
    if kind == "os2":
        return np.concatenate([np.asarray(o[0], float), np.asarray(o[1], float)])
    return np.asarray(o, float).reshape(-1)


PLANS = {
    "cnt0": [(0, "zero"), (1, "up"), (2, "walk"), (3, "random")],
    "cnt1": [(0, "zero"), (1, "up"), (2, "walk"), (3, "random")],
    "cnt2": [(0, "zero"), (1, "up"), (2, "walk"), (3, "dose")],
    "ma1": [(0, "up"), (1, "walk"), (2, "random"), (3, "dose")],
    "os2": [(0, "const"), (1, "walk"), (2, "lowdose"), (3, "dose")],
}


def make_cnt(out_dir, versions):
    for kind, plans in PLANS.items():
        for seed, plan in plans:
            ep = run_episode(kind, seed, plan)
            print("%-4s seed %d %-7s steps %3d sumR %9.4g Vmax %8.4g physical %-5s finite %s warn %d" % (
                kind, seed, plan, ep["n_steps"], ep["reward"].sum(), ep["v_max"], ep["physical"],
                bool(np.isfinite(ep["obs"]).all()), ep["odeint_warnings"]), flush=True)
            np.savez_compressed(os.path.join(out_dir, "cnt_%s_seed%d_%s.npz" % (kind, seed, plan)), versions=versions,
                                **ep)


if __name__ == "__main__":
    import scipy
    out = os.path.join(os.path.dirname(HERE), "tests", "golden")
    make_cnt(out, np.array([np.__version__, scipy.__version__, sys.version.split()[0]]))
