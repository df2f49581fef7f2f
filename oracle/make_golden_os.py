"""Golden episodes of the UNMODIFIED reference `SBROS-v1` env (gym_SBR_oneshot.py), authoring container only.
Called from oracle/make_golden.py.  Each episode: np.random.seed(seed); reset(); step(action_k) until done."""
import os
import warnings

import numpy as np

import ref_shim


def action_plan(kind, rng, n=600):
    if kind == "const":
        return np.tile([2.0, 5.0], (n, 1))
    if kind == "const_hi":
        return np.tile([4.0, 8.0], (n, 1))
    if kind == "clip":                        # outside the clip ranges [0,8] x [0,15] (gym_SBR_oneshot.py:865-906)
        return np.tile([9.5, -3.0], (n, 1))
    if kind == "random":
        return np.stack([8 * rng.rand(n), 15 * rng.rand(n)], axis=1)
    if kind == "walk":                        # slowly varying set-points, as an RL policy would produce
        a = np.cumsum(0.15 * rng.randn(n, 2), axis=0) + [2.0, 6.0]
        return np.stack([np.clip(a[:, 0], 0.5, 7), np.clip(a[:, 1], 0, 14)], axis=1)
    if kind == "aggr":                        # aggressive but physical: DO set-point slammed between the ends of its
        do = np.where((np.arange(n) // 10) % 2 == 0, 7.5, 0.5)             # range every 10 steps, NO3 set-point
        return np.stack([do, 9 + 6 * rng.rand(n)], axis=1)                 # high enough that little carbon is dosed
    if kind == "aggr_random":                 # per-step random set-points over the upper NO3 range
        return np.stack([8 * rng.rand(n), 8 + 7 * rng.rand(n)], axis=1)
    raise ValueError(kind)


def run_episode(seed, kind, tight=None):
    import gym_SBR.envs.gym_SBR_oneshot as m
    rng = np.random.RandomState(1000 + seed)
    plan = action_plan(kind, rng)
    np.random.seed(seed)
    env = m.SbrOS()
    rec = dict(obs_do=[], obs_ec=[], state=[], reward=[], done=[], action=[], t=[], warn=[])
    with ref_shim.quiet(), warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        obs0 = env.reset()
        influent = np.array(m.influent_mixed, dtype=float)
        x_fill = np.array(m.x_out[-1], dtype=float)
        k = 0
        while True:
            obs, state, reward, done, info = env.step(plan[k])
            rec["obs_do"].append(obs[0]); rec["obs_ec"].append(obs[1]); rec["state"].append(np.array(state))
            rec["reward"].append(float(reward)); rec["done"].append(bool(done)); rec["action"].append(plan[k])
            rec["t"].append(float(m.t))
            rec["warn"].append(len(wlist))        # cumulative count of scipy ODEintWarnings (LSODA gave up)
            k += 1
            if done or k >= 600:
                break
    out = {k_: np.array(v) for k_, v in rec.items()}
    out.update(seed=seed, kind=kind, reset_obs_do=np.array(obs0[0]), reset_obs_ec=np.array(obs0[1]),
               influent=influent, x_fill=x_fill, n_steps=k, Qw=float(m.Qw),
               kla_list=np.array(m.Kla, dtype=float), ec_tail=np.array(m.EC[-2000:], dtype=float),
               so_tail=np.array(m.So[-600:], dtype=float), ie_DO_last=float(m.ie_DO[-1]), ie_EC_last=float(m.ie_EC[-1]))
    return out


def make_os(out_dir, versions, only=None):
    from make_golden import tight_odeint
    episodes = [(0, "const"), (1, "const_hi"), (2, "walk"), (3, "clip"), (4, "random"), (5, "walk"), (6, "walk"),
                (7, "aggr"), (8, "aggr_random")]
    if only is not None:
        episodes = [e for e in episodes if e in only]
    for seed, kind in episodes:
        try:
            ep = run_episode(seed, kind)
        except Exception as e:                      # the reference itself can fail under hard actions
            print("os seed %d %s: reference raised %r -- skipped" % (seed, kind, e), flush=True)
            continue
        print("os seed %d %-8s steps %d sumR %.9g finite %s" % (seed, kind, ep["n_steps"], ep["reward"].sum(),
                                                               bool(np.isfinite(ep["state"]).all())), flush=True)
        np.savez_compressed(os.path.join(out_dir, "sbros_v1_seed%d_%s.npz" % (seed, kind)), versions=versions, **ep)
    if only is not None:
        return
    with tight_odeint():
        ep = run_episode(0, "const")
    print("os tight seed 0 const sumR %.9g" % ep["reward"].sum(), flush=True)
    np.savez_compressed(os.path.join(out_dir, "sbros_v1_seed0_const_tight.npz"), versions=versions, **ep)
