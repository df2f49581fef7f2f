"""CPU checks of the stepper logic behind `SBRCnt-v0/1/2`, `SBRCntMA-v1`, `SBROS-v2` (csrc/sbr_cnt.cuh compiled with
g++) against whole episodes of the UNMODIFIED reference env modules run with their shared reward function repaired
(oracle/make_golden_cnt.py holds the disclosure; fixtures tests/golden/cnt_*.npz).

What is compared, step by step: the 14-component reactor state the next step continues from, the running time, both
set-points, KLa and the dosing flow, the observation the env returns, `done`, and the (repaired) reward.  Episodes in
which the reference's unclamped carbon controller runs away (`physical` False: reactor volume up to hundreds of m3)
are followed up to the last reacting step; their settle/draw step decants a number of layers the reference's slice
arithmetic was never meant for (status SBR_ST_LAYERS here)."""
import glob
import os

import numpy as np
import pytest

from gym_sbr2_b200 import _abi, cnt, parity
from oracle.twin import binding as twin

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CNT_EPISODES = sorted(os.path.basename(f)[4:-4] for f in glob.glob(os.path.join(GOLDEN, "cnt_*.npz")))
STATE_ATOL = 1e-7           # as for the other interval-per-step paths (parity.OS_ATOL): 466 chained LSODA restarts
# thresholds of the repaired reward (module_reward_continuous1.py:32-41): a So within tolerance of one may flip the bin
R_EDGES = (1.5, 2.5, 3.5, 5.0)


def load_cnt(kind, name):
    return np.load(os.path.join(GOLDEN, "cnt_%s_%s.npz" % (kind, name)), allow_pickle=True)


def obs_close(obs, ref, kind):
    """Observations are states (or step-deltas of states) under O(1) scales; cnt0 divides Snh by 0.005 and So by 2."""
    obs, ref = np.asarray(obs, float)[:len(ref)], np.asarray(ref, float)
    atol = np.full(len(ref), 1e-6)
    if kind == "cnt0":
        atol[6] = STATE_ATOL * 20 / 0.005
    ratio = np.abs(obs - ref) / (parity.RTOL * np.abs(ref) + atol)
    return bool(np.all(np.isfinite(obs)) and ratio.max() <= 1.0), float(ratio.max())


def run_cnt(make_batch, episode):
    kind, name = episode.split("_", 1)
    g = load_cnt(kind, name)
    b = make_batch(kind)
    ob0 = b.reset(g["influent"][:, None])
    n = int(g["n_steps"])
    last = n if bool(g["physical"]) else n - 1
    assert obs_close(ob0[:, 0], g["reset_obs"], kind)[0]
    assert parity.state_close(b.st[:14, 0], g["x_fill"], atol_frac=STATE_ATOL)[0]
    worst = 0.0
    for k in range(last):
        a = np.zeros((2, 1))
        a[:, 0] = g["action"][k]
        ob, r, d = b.step(a)
        assert b.status[0] == 0, (episode, k, int(b.status[0]))
        ok, w = parity.state_close(b.st[:14, 0], g["x_cont"][k], atol_frac=STATE_ATOL)
        assert ok, (episode, k, w)
        worst = max(worst, w)
        ok, w = obs_close(ob[:, 0], g["obs"][k], kind)
        assert ok, (episode, k, w)
        assert bool(d[0]) == bool(g["done"][k]), (episode, k)
        if k < n - 1:
            assert abs(b.st[_abi.CNT_T, 0] - g["t"][k]) < 1e-12, (episode, k)
        assert abs(b.st[_abi.CNT_U_DO, 0] - g["u_do"][k]) < 1e-12 and abs(b.st[_abi.CNT_U_EC, 0] - g["u_ec"][k]) < 1e-12
        # KLa = Kc_DO * (set-point - So) + ...: the state tolerance on So (1e-5 * So + 8e-7) times the gain
        assert abs(b.st[_abi.CNT_KLA_LAST, 0] - g["kla"][k]) <= 1e-5 * abs(g["kla"][k]) + b.cfg.Kc_DO * 2e-5, (episode, k)
        # the dosing flow is an INTEGRATING controller output, EC[k] = EC[k-1] + Kc_EC * (set-point - measured value) + ...:
        # it inherits the state tolerance times the gain, summed over the intervals run so far
        assert abs(b.st[_abi.CNT_EC_LAST, 0] - g["ec"][k]) <= 1e-5 * abs(g["ec"][k]) + b.cfg.Kc_EC * 4e-6 * (k + 1), \
            (episode, k)
        so = g["x_end"][k][8]
        if min(abs(so - e) for e in R_EDGES) > 1e-4 or bool(g["done"][k]):
            assert r[0] == g["reward"][k], (episode, k, r[0], g["reward"][k])
    if last == n:
        assert b.done[0] == 1 and b.st[_abi.CNT_STEPS, 0] == n == cnt.EPISODE_STEPS[kind]
        assert abs(b.st[_abi.CNT_QW, 0] - float(g["Qw"])) <= 1e-5 * abs(float(g["Qw"])), episode
    return worst


@pytest.mark.parametrize("episode", CNT_EPISODES)
def test_dp45_episodes_match_reference(built, episode):
    run_cnt(lambda kind: twin.CntBatch(cnt.cnt_config(kind), 1, cnt.OBS_ROWS[kind], mode=_abi.MODE_DP45), episode)


def test_fixture_coverage():
    """Every kind has at least one episode the reference keeps physical from reset to done and, where the env has a
    carbon controller, at least one in which that controller doses."""
    for kind in cnt.KINDS:
        eps = [load_cnt(*e.split("_", 1)) for e in CNT_EPISODES if e.startswith(kind + "_")]
        assert len(eps) >= 4 and any(bool(g["physical"]) for g in eps), kind
        if kind in ("cnt2", "ma1", "os2"):
            assert any(float(np.max(g["ec"])) > 0 for g in eps), kind


def test_finished_episode_is_a_noop(built):
    g = load_cnt("cnt1", "seed0_zero")
    b = twin.CntBatch(cnt.cnt_config("cnt1"), 1, 5, mode=_abi.MODE_DP45)
    b.reset(g["influent"][:, None])
    for k in range(int(g["n_steps"])):
        b.step(np.array([[g["action"][k][0]], [0.0]]))
    assert b.done[0] == 1
    before = b.st.copy()
    _, r, _ = b.step(np.array([[0.4], [0.0]]))
    assert np.array_equal(b.st, before, equal_nan=True) and r[0] == 0.0 and b.status[0] == _abi.ST_DONE


def test_rk4_fine_grid_agrees_with_dp45(built):
    """The fixed-step mode on a fine grid (20 sub-steps per control interval) and the adaptive mode are two independent
    discretisations of the same episode."""
    from gym_sbr2_b200 import schedule
    g = load_cnt("ma1", "seed0_up")
    a = twin.CntBatch(cnt.cnt_config("ma1"), 1, 5, mode=_abi.MODE_DP45)
    b = twin.CntBatch(cnt.cnt_config("ma1"), 1, 5, mode=_abi.MODE_RK4,
                      sched=schedule.os_schedule(rk4_sub_interval=20, rk4_sub_fill=2000, rk4_sub_idle=3000))
    a.reset(g["influent"][:, None]); b.reset(g["influent"][:, None])
    for k in range(200):
        act = np.array([[g["action"][k][0]], [0.0]])
        a.step(act); b.step(act)
        assert parity.state_close(a.st[:14, 0], b.st[:14, 0], atol_frac=STATE_ATOL)[0], k


def random_plan(kind, rng, n):
    """Action sequences that keep the reference physical: DO set-points moving around 1-3 g/m3, carbon set-point at 0."""
    a = np.zeros((n, 2))
    if kind == "os2":
        a[:, 0] = np.clip(2.0 + np.cumsum(0.1 * rng.randn(n)), 0.3, 5.0)
        return a
    a[:, 0] = 0.03 * rng.randn(n) * (0.1 if kind == "cnt0" else 1.0)
    up = range(60, 64) if kind == "ma1" else range(1, 5)
    for k in up:
        a[k, 0] += rng.uniform(0.25, 0.75) * (0.06 if kind == "cnt0" else 1.0)
    if kind in ("cnt2", "ma1"):
        a[0, 0] = -2.0
    if kind == "ma1":
        a[1:60, 0] = 0.0
    return a


@pytest.mark.parametrize("kind", sorted(cnt.KINDS))
def test_random_actions_against_oracle(built, kind):
    """Fresh influent draws and random action sequences (not the fixtures): the g++ build of the stepper against the scipy
    restatement of the reference env (oracle.SbrCntOracle, itself pinned to the reference by test_oracle_golden_cnt)."""
    import warnings
    from gym_sbr2_b200 import influent, schedule
    from oracle import sbr_oracle as O
    rng = np.random.RandomState(77)
    steps = 140 if kind in ("cnt0", "ma1", "os2") else 60
    for trial in range(2):
        infl = influent.mix_numpy(0, rng.randn(48))
        plan = random_plan(kind, rng, steps)
        o = O.SbrCntOracle(kind)
        b = twin.CntBatch(cnt.cnt_config(kind), 1, cnt.OBS_ROWS[kind], mode=_abi.MODE_DP45)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o.reset(np.concatenate([[0.66], infl[1:]]))
            load = infl.copy()
            load[0] = schedule.os_fill_flow(b.params.Qin)          # influent_mixed[0] := Qin / t_memory1[-1]
            b.reset(load[:, None])
            for k in range(steps):
                out = o.step(plan[k] if kind == "os2" else plan[k][:1])
                ob, r, d = b.step(plan[k][:, None])
                ok, w = parity.state_close(b.st[:14, 0], o.x, atol_frac=STATE_ATOL)
                assert ok, (kind, trial, k, w)
                ref_obs = np.concatenate(out[0]) if kind == "os2" else np.asarray(out[0]).reshape(-1)
                assert obs_close(ob[:, 0], ref_obs, kind)[0], (kind, trial, k)
                assert float(np.max(o.x[0])) < 1.4
