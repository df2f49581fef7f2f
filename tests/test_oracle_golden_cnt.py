"""The oracle's restatement of `SbrCnt0/1/2`, `SbrCntMA1`, `SbrOS1` (oracle/sbr_oracle.py: SbrCntOracle) against whole
episodes of the unmodified reference env modules (reward repaired as oracle/make_golden_cnt.py discloses).

SbrCnt0 / SbrCnt1 are reproduced BIT FOR BIT (same odeint calls on the same grids with the same RHS arithmetic).  The
three envs with a carbon controller are reproduced to ~1e-8 relative: their reference fill RHS rewrites LSODA's state
array in place on every call (`x[i] = x[i] * x[0] / (x[0] + ec)` with ec = 0, gym_SBR_continuous2.py:634-662), a no-op up
to one rounding of x * V / V that the restatement does not imitate."""
import warnings

import numpy as np
import pytest

from oracle import sbr_oracle as O
from test_twin_parity_cnt import CNT_EPISODES, load_cnt


def replay(kind, g, steps=None):
    o = O.SbrCntOracle(kind)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                      # LSODA's "excess work" chatter on run-away episodes
        ob0 = o.reset(np.concatenate([[0.66], g["influent"][1:]]))
        ob0 = np.concatenate(ob0) if kind == "os2" else np.asarray(ob0).reshape(-1)
        rows = []
        for k in range(steps or int(g["n_steps"])):
            out = o.step(g["action"][k] if kind == "os2" else g["action"][k][:1])
            if kind == "os2":
                (od, oe), st, r, d = out
                ob = np.concatenate([od, oe])
            else:
                ob, r, d = out
                ob, st = np.asarray(ob).reshape(-1), np.zeros(0)
            rows.append((o.x.copy(), ob, st, r, bool(d), o.t, o.u_do, o.u_ec, o.Kla[-1], o.EC[-1]))
    return ob0, rows


@pytest.mark.parametrize("episode", CNT_EPISODES)
def test_oracle_reproduces_reference_episodes(episode):
    kind, name = episode.split("_", 1)
    g = load_cnt(kind, name)
    n = int(g["n_steps"])
    ob0, rows = replay(kind, g)
    exact = kind in ("cnt0", "cnt1")
    rtol = 0.0 if exact else 1e-6          # 1e-8-level LSODA path differences, amplified in small per-step deltas
    close = lambda a, b: np.all(np.abs(np.asarray(a) - np.asarray(b)) <= rtol * np.abs(b) + (0.0 if exact else 1e-8))
    assert close(ob0, g["reset_obs"])
    last = n if bool(g["physical"]) else n - 1
    for k in range(last):
        x, ob, st, r, d, t, u_do, u_ec, kla, ec = rows[k]
        assert close(x, g["x_cont"][k]), (episode, k, np.abs(x - g["x_cont"][k]).max())
        assert close(ob, g["obs"][k]), (episode, k)
        if kind == "os2":
            assert close(st, g["state15"][k]), (episode, k)
        assert r == g["reward"][k] and d == bool(g["done"][k]), (episode, k)
        if k < n - 1:
            assert t == g["t"][k]
        assert u_do == g["u_do"][k] and u_ec == g["u_ec"][k]
        if exact:
            assert kla == g["kla"][k] and ec == g["ec"][k], (episode, k)
        else:       # controller outputs carry the state difference times the gain (and the carbon controller integrates it)
            assert abs(kla - g["kla"][k]) <= 1e-6 * abs(g["kla"][k]) + 1e-5, (episode, k)
            assert abs(ec - g["ec"][k]) <= 1e-6 * abs(g["ec"][k]) + 1e-6 * (k + 1), (episode, k)
    assert rows[n - 1][4] is True and not any(r[4] for r in rows[:n - 1])


def test_episode_lengths_and_phase_stamps():
    tm = O.batch_time_stamps()
    assert [len(t) for t in tm[:5]] == [25, 46, 190, 171, 1]          # t_memory1..5 of module_batch_time at t_delta = 10 dt
    assert abs(tm[1][0] - 0.021833333333333333) < 1e-15 and abs(tm[3][-1] - 0.4085000000000001) < 1e-15
    for kind, steps in (("cnt0", 466), ("cnt1", 228), ("cnt2", 228), ("ma1", 463), ("os2", 463)):
        g = load_cnt(kind, {"cnt0": "seed0_zero", "cnt1": "seed0_zero", "cnt2": "seed0_zero", "ma1": "seed0_up",
                            "os2": "seed0_const"}[kind])
        assert int(g["n_steps"]) == steps
