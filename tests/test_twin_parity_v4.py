"""CPU checks of the SBR-v4 stepper logic the CUDA kernel inlines (csrc/sbr_core.cuh compiled with g++) against whole
episodes of the reference `SbrEnv4` (numpy < 1.18 linspace semantics, see test_oracle_golden_v4.py) and against the
oracle run with LSODA at 1e-12."""
import numpy as np
import pytest

from gym_sbr2_b200 import _abi, parity
from oracle import sbr_oracle as O
from oracle.twin import binding as twin
from test_oracle_golden_v4 import V4_EPISODES, load_v4

# The reference's own LSODA (rtol = atol = 1.49e-8, restarted every 72 s under a PID with derivative action
# 300 * dSo) ends up to 1.9 tolerance units from LSODA at 1e-12 in So on these episodes (measured; every other
# component <= 0.11), while this code's DP45 path stays below 0.05 units on all of them.  Against the default
# reference the bound on So is therefore 4 units; against the tight oracle the plain tolerance applies.
V4_SO_SLACK = 4.0


def v4_state_close(obs, ref_state, so_slack=1.0, atol_frac=parity.OS_ATOL):
    raw, ref = np.asarray(obs) * O.X1_V4, np.asarray(ref_state) * O.X1_V4
    bound = parity.RTOL * np.abs(ref) + atol_frac * parity.STATE_SCALE
    bound[8] *= so_slack
    ratio = np.abs(raw - ref) / bound
    return bool(np.all(np.isfinite(raw)) and ratio.max() <= 1.0), float(ratio.max())


def run_v4(make_batch, names, steps=493):
    G = [load_v4(nm) for nm in names]
    b = make_batch(len(G))
    ob0 = b.reset(np.stack([g["influent"] for g in G], axis=1))
    rec = []
    for k in range(steps):
        out = b.step(np.array([g["action"][k] for g in G]))
        rec.append(out + (b.status.copy(), b.counters.copy()))
    return G, b, ob0, rec


def check_v4(G, b, ob0, rec, names):
    for j, g in enumerate(G):
        assert np.allclose(ob0[:, j], g["reset_obs"], rtol=1e-14, atol=0), names[j]
        for k in range(len(rec)):
            ob, r, done, status, _ = rec[k]
            assert bool(done[j]) == bool(g["done"][k]), (names[j], k)
            assert status[j] == 0, (names[j], k)
            ok, worst = v4_state_close(ob[:, j], g["state"][k], so_slack=V4_SO_SLACK)
            assert ok, (names[j], k, worst)
            # rewards are 0.5 - O(1e-3); the terminal one carries the +-246 ammonia step (no golden sits on it)
            assert abs(r[j] - g["reward"][k]) <= 1e-5 * abs(g["reward"][k]) + 1e-9, (names[j], k)
        if len(rec) == int(g["n_steps"]):
            assert abs(g["eff"][3] - 4) > 1e-2
            assert abs(b.st[_abi.V4_QW, j] - float(g["Qw"])) <= 1e-5 * float(g["Qw"]), names[j]
            assert b.st[_abi.V4_STEPS, j] == 493
            assert abs(b.st[_abi.V4_KLA_SUM, j] - float(g["kla_sum"])) <= 1e-5 * float(g["kla_sum"])


def test_dp45_episodes_match_reference(built):
    G, b, ob0, rec = run_v4(lambda n: twin.V4Batch(n, mode=_abi.MODE_DP45), V4_EPISODES)
    check_v4(G, b, ob0, rec, V4_EPISODES)


@pytest.mark.parametrize("name", ["seed4_walk", "seed3_random"])
def test_dp45_against_tight_oracle(built, name):
    g = load_v4(name)
    b = twin.V4Batch(1, mode=_abi.MODE_DP45)
    b.reset(g["influent"][:, None])
    o = O.SbrEnv4Oracle(ode_kw=dict(rtol=1e-12, atol=1e-12, mxstep=50000))
    o.reset(np.concatenate([[0.66], g["influent"][1:]]))
    for k in range(493):
        ob, r, done = b.step(np.array([g["action"][k]]))
        st, rt, dt = o.step(float(g["action"][k]))
        ok, worst = v4_state_close(ob[:, 0], st, atol_frac=1e-8)
        assert ok, (k, worst)
        assert abs(r[0] - rt) <= 1e-6 * abs(rt) + 1e-10 and bool(done[0]) == dt


def test_phase_sequence_and_counters(built):
    """26 fill steps, 466 react steps, then ONE step that settles, draws and idles; RK4 on the reference grid makes
    the 9/10 output-point pattern and the 361-point idle solve visible in the RHS counters."""
    g = load_v4("seed0_zero")
    b = twin.V4Batch(1, mode=_abi.MODE_RK4)
    b.reset(g["influent"][:, None])
    rhs = []
    for k in range(493):
        b.step(np.array([g["action"][k]]))
        rhs.append(int(b.counters[0, 0]))
    assert set(rhs[:-1]) <= {32, 36}
    assert rhs[-1] % 4 == 0 and rhs[-1] > 4 * 300
    assert list(b.done) == [1]
    before = b.st.copy()
    ob, r, d = b.step(np.array([0.3]))                         # finished episode: no-op
    assert np.array_equal(b.st, before, equal_nan=True) and r[0] == 0.0 and b.status[0] == _abi.ST_DONE
