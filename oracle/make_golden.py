"""Generate tests/golden/*.npz by running the UNMODIFIED reference (authoring container only).

    python oracle/make_golden.py [v2] [os] [rhs]

Test infrastructure: imports /root/reference through oracle/ref_shim.py (gym/matplotlib stubbed, numerics
untouched) and records inputs + outputs of the two envs that step on a current toolchain:
`SBR-v2` (cycle-per-step) and `SBROS-v1` (interval-per-step), plus stage-level samples (RHS values, settle,
draw).  A second set is produced with scipy's odeint forced to rtol=atol=1e-12 ("tight") so the tests can
separate the reference's own integration error (~2.6e-6) from ours.  The fixtures record numpy/scipy
versions.  /root/reference does not exist on the GPU box; the fixtures are what travels.
"""
import os
import sys

import numpy as np
import scipy
import scipy.integrate

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
VERSIONS = np.array([np.__version__, scipy.__version__, sys.version.split()[0]])

_real_odeint = scipy.integrate.odeint


class tight_odeint(object):
    """Context manager: every odeint call made by the reference uses rtol=atol=tol."""

    def __init__(self, tol=1e-12):
        self.tol = tol

    def __enter__(self):
        tol = self.tol

        def od(func, y0, t, args=(), **kw):
            kw.setdefault("rtol", tol)
            kw.setdefault("atol", tol)
            kw.setdefault("mxstep", 50000)
            return _real_odeint(func, y0, t, args=args, **kw)

        scipy.integrate.odeint = od
        return self

    def __exit__(self, *a):
        scipy.integrate.odeint = _real_odeint


V2_ACTIONS = [
    [0.25, 0.25, 0.25], [0.25, 0.5, 0.75], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [-0.5, 1.5, 0.3],
    [0.05, 0.9, 0.1], [0.6, 0.02, 0.4],
]


def run_v2_case(seed, action, tight=False):
    """reset + one step of the reference SbrEnv2 with np.random.seed(seed); capture SBR_model_FB.run outputs."""
    import gym_SBR.envs.gym_SBR_env2 as m
    np.random.seed(seed)
    env = m.SbrEnv2()
    cap = {}
    orig = env._next_observation

    def spy(*a, **k):
        out = orig(*a, **k)
        cap["run"] = out
        return out

    env._next_observation = spy
    with ref_shim.quiet():
        obs0 = env.reset()
        influent_reset = np.array(m.influent_mixed, dtype=float)
        if tight:
            with tight_odeint():
                obs, reward, done, info = env.step(np.array(action, dtype=float))
        else:
            obs, reward, done, info = env.step(np.array(action, dtype=float))
    (t, x, x_last, sp3, So3, t3, sp5, So5, t5, sp8, So8, t8, Qeff, eff, Qw, kla3, kla5, kla8, EQI) = cap["run"]
    return dict(seed=seed, action=np.array(action, dtype=float), reset_obs=np.array(obs0, dtype=float),
                influent=influent_reset, influent_step=np.array(m.influent_mixed, dtype=float),
                obs=np.array(obs, dtype=float), reward=float(reward), done=bool(done),
                x_last=np.array(x_last, dtype=float), Qw=float(Qw), EQI=float(EQI),
                eff=np.array(eff, dtype=float), Qeff=float(Qeff),
                kla3_mean=float(np.mean(kla3)), kla5_mean=float(np.mean(kla5)), kla8_mean=float(np.mean(kla8)),
                kla3_last=float(kla3[-1]), kla5_last=float(kla5[-1]), kla8_last=float(kla8[-1]),
                n3=len(kla3), n5=len(kla5), n8=len(kla8), n_traj=len(t))


def stack(cases):
    keys = cases[0].keys()
    return {k: np.array([c[k] for c in cases]) for k in keys}


def make_v2():
    cases, tight = [], []
    for seed in range(6):
        for ai, action in enumerate(V2_ACTIONS):
            if seed >= 2 and ai >= 2 and (seed + ai) % 3:
                continue
            cases.append(run_v2_case(seed, action))
            print("v2 seed %d action %s reward %.15g" % (seed, action, cases[-1]["reward"]), flush=True)
    rng = np.random.RandomState(1234)
    for seed in range(6, 18):
        cases.append(run_v2_case(seed, rng.rand(3).tolist()))
        print("v2 seed %d random reward %.15g" % (seed, cases[-1]["reward"]), flush=True)
    for seed, action in [(0, V2_ACTIONS[0]), (0, V2_ACTIONS[1]), (1, V2_ACTIONS[5]), (3, V2_ACTIONS[3])]:
        tight.append(run_v2_case(seed, action, tight=True))
        print("v2 tight seed %d reward %.15g" % (seed, tight[-1]["reward"]), flush=True)
    np.savez(os.path.join(OUT, "sbr_v2_cases.npz"), versions=VERSIONS, **stack(cases))
    np.savez(os.path.join(OUT, "sbr_v2_tight.npz"), versions=VERSIONS, **stack(tight))


def make_rhs():
    """Stage-level samples straight from the reference's classes."""
    import gym_SBR.envs.sub_phases_FB as sp
    import gym_SBR.envs.gym_SBR_oneshot as os_mod
    import gym_SBR.envs.SBR_model_FB  # noqa: F401
    rng = np.random.RandomState(7)
    Spar = [0.24, 0.67, 0.08, 0.08, 0.06]
    Kpar = [4.0, 10.0, 0.2, 0.5, 0.3, 0.8, 0.8, 3.0, 0.1, 0.5, 1.0, 0.05, 0.4, 0.05]
    so_sat = float(os_mod.DO_set(15))
    dcp = [5.0, 0.00035, 0.02 / 24, 2, 0, 240, 12, 2, 5, 0.005, so_sat]
    x0 = np.array([0.6161484733495801, 30, 0.571098000538576, 1440.01157895393, 31.254221999137, 2599.2714348941,
                   168.915006750837, 551.901552960823, 2.16607843793004, 13.3791460027604, 0.00562880208518134,
                   0.35996687629947, 1.86916737961228, 3.790463057094611])
    load = np.array([33.51673936430571, 30.0, 69.8, 41.7, 175.7, 24.1, 0, 0, 0, 0, 33.2, 6.98, 9.08, 7.0])
    X, KLA, EC, D_react, D_fill, D_ec = [], [], [], [], [], []
    fill, rxn, osenv = sp.filling(1.32, load[0]), sp.rxn(1.32), os_mod.SbrOS()
    for i in range(48):
        x = x0 * np.exp(0.6 * rng.randn(14))
        if i % 5 == 0:
            x[8] = 1e-9 * rng.rand()
        kla = 240 * rng.rand()
        ec = 0.0005 * rng.rand() * (i % 2)
        X.append(x)
        KLA.append(kla)
        EC.append(ec)
        D_react.append(rxn.dxdt(x.copy(), 0.0, Spar, Kpar, dcp, kla))
        D_fill.append(fill.dxdt(x.copy(), 0.0, Spar, Kpar, dcp, kla, load))
        D_ec.append(osenv.reaction_dxdt(x.copy(), 0.0, Spar, Kpar, dcp, None, kla, ec))
    # settle + draw on a handful of end-of-react states
    SX, XF, X7, QW, EQI, EFF, XS = [], [], [], [], [], [], []
    with ref_shim.quiet():
        for i in range(8):
            x = x0 * np.exp(0.05 * rng.randn(14))
            x[0] = 1.32 * (1 + 0.002 * rng.randn())
            t0, t1 = 0.4169166666666667, 0.4169166666666667 + 0.5 * 0.083
            _, _, sX, Xf = sp.settling().sim_settling(t0, t1, 0.002 / 24, x.copy())
            _, x7, Qw, _, _, eqi, eff = sp.drawing().sim_drawing(0, 0.0105, 0.002 / 24, x.copy(), sX.copy(), Xf,
                                                                 0.66, 2700)
            XS.append(x)
            SX.append(sX)
            XF.append(Xf)
            X7.append(x7)
            QW.append(Qw)
            EQI.append(eqi)
            EFF.append(eff)
    np.savez(os.path.join(OUT, "stage_samples.npz"), versions=VERSIONS, x=np.array(X), kla=np.array(KLA),
             ec=np.array(EC), load=load, ec_conc=float(os_mod.EC_conc), so_sat=so_sat,
             d_react=np.array(D_react), d_fill=np.array(D_fill), d_ec=np.array(D_ec),
             settle_x=np.array(XS), settle_sX=np.array(SX), settle_Xf=np.array(XF), draw_x7=np.array(X7),
             draw_Qw=np.array(QW), draw_EQI=np.array(EQI), draw_eff=np.array(EFF, dtype=float))


if __name__ == "__main__":
    what = sys.argv[1:] or ["v2", "rhs", "os"]
    os.makedirs(OUT, exist_ok=True)
    ref_shim.load_reference()
    if "rhs" in what:
        make_rhs()
    if "v2" in what:
        make_v2()
    if "os" in what:
        from make_golden_os import make_os
        make_os(OUT, VERSIONS)
