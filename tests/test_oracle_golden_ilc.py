"""Pin the ILC oracle (oracle/sbr_oracle_ilc.py) to outputs of the reference's own batch-to-batch feed-forward path
(module_batch_PID.batch_PID, SBR_model_PID_on.run, SBR_model_batchPID_fbPID.run) recorded by oracle/make_golden_ilc.py.

DISCLOSURE: sub_phases_batchPID_fbPID.py needs numpy < 1.18 linspace semantics (float `num` -> int(num)), restored for that
module only; the reward of `SBR-v0` cannot be computed by the reference at all (seven arguments into a ten-parameter
function, gym_SBR_env0.py:203) and is not part of the fixtures.  See the header of oracle/make_golden_ilc.py."""
import os

import numpy as np
import pytest

from oracle import sbr_oracle as O
from oracle import sbr_oracle_ilc as I

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ilc_seed%d.npz")
NAMES = ("1", "2", "3", "4", "5", "8")


@pytest.fixture(scope="module", params=[0, 1], ids=["seed0", "seed1_box_edges"])
def g(request):
    """Fixture 0: mid-range set-points; fixture 1: another influent draw and set-points at the edges of the action box."""
    return np.load(GOLDEN % request.param, allow_pickle=True)


def test_fixture_inventory(g):
    assert [len(g["t_memory" + n]) for n in NAMES] == [217, 433, 2008, 1675, 111, 325]
    assert g["actions"].shape == (3, 3) and g["actions_learn"].shape == (3, 3) and g["influent"][0] == I.FILL_FLOW_ILC
    assert np.array_equal(g["x0"], np.array(I.X0_ILC)) and np.array_equal(g["par_batchPID"], np.array(I.PAR_BATCH_PID))
    # the stamp lists are module_batch_time's (already restated for the SBRCnt family)
    stamps = O.batch_time_stamps(t_delta=O.DT)
    for n, k in zip(NAMES, I.PHASES):
        assert np.array_equal(np.array(stamps[k]), g["t_memory" + n])


def test_cycle0_feed_forward_base_matches_reference(g):
    """SBR_model_PID_on.run at import (gym_SBR_env0.py:105-106): x_last, per-sample So and KLa of the six phases."""
    r = I.ilc_cycle(g["x0"], g["influent"], [0, 0, 2, 0, 2, 0, 0, 2])
    assert np.allclose(r["x_last"], g["x_last0"], rtol=1e-9, atol=1e-12)
    for j, n in enumerate(NAMES):
        assert len(r["So_memory"][j]) == len(g["So0_" + n])
        assert np.allclose(r["So_memory"][j], g["So0_" + n], rtol=1e-8, atol=1e-12), n
        assert np.allclose(r["Kla_memory"][j], g["kla0_" + n], rtol=1e-8, atol=1e-10), n


def _weights(g):
    return I.ilc_weights([g["t_memory" + n] if k in I.PHASES else None
                          for k, n in zip(range(8), ("1", "2", "3", "4", "5", "x", "x", "8"))])


@pytest.mark.parametrize("chain", ["env", "learn"])
def test_batch_pid_matches_reference(g, chain):
    """E_batch and u_batch of three consecutive cycles; `env` = the memories frozen at cycle 0 (what SbrEnv.step does),
    `learn` = the previous cycle's memories fed back."""
    wt = _weights(g)
    mem = I.IlcMemory([len(g["t_memory" + n]) for n in NAMES])
    so = [g["So0_" + n] for n in NAMES]
    sp = [g["sp0_" + n] for n in NAMES]
    for c, a in enumerate(g["actions" if chain == "env" else "actions_learn"]):
        sp_in = list(sp)
        for j, av in ((2, a[0]), (4, a[1]), (5, a[2])):
            sp_in[j] = I.ilc_setpoint_memory(sp[j], av)
        E = [I.ilc_e_batch(sp_in[j], so[j], wt[j][0], wt[j][1]) for j in range(6)]
        u = mem.update(E)
        for j, n in enumerate(NAMES):
            assert np.allclose(E[j], g["%s_c%d_E%s" % (chain, c, n)], rtol=1e-12, atol=1e-15), (c, n)
            assert np.allclose(u[j], g["%s_c%d_u%s" % (chain, c, n)], rtol=1e-12, atol=1e-14), (c, n)
        if chain == "learn":
            so = [g["learn_c%d_So%s" % (c, n)] for n in NAMES]
            sp = [np.full(len(so[j]), v) for j, v in enumerate((0, 0, a[0], 0, a[1], a[2]))]


@pytest.mark.parametrize("c", [0, 1, 2])
def test_feed_forward_cycle_matches_reference(g, c):
    """SBR_model_batchPID_fbPID.run with the reference's own u_batch rows: end state, per-sample So, the clamped
    feed-forward profile, Qeff / Qw."""
    a = g["actions_learn"][c]
    x_in = g["x_last0"] if c == 0 else g["learn_c%d_x_last" % (c - 1)]
    r = I.ilc_cycle(x_in, g["influent"], [0, 0, a[0], 0, a[1], 0, 0, a[2]],
                    kla_memory=[g["kla0_" + n] for n in NAMES], u_batch=[g["learn_c%d_u%s" % (c, n)] for n in NAMES])
    assert np.allclose(r["x_last"], g["learn_c%d_x_last" % c], rtol=1e-9, atol=1e-12)
    assert np.allclose([r["Qeff"], r["Qw"]], g["learn_c%d_Qeff_Qw" % c], rtol=1e-9, atol=1e-13)
    for j, n in enumerate(NAMES):
        assert np.allclose(r["So_memory"][j], g["learn_c%d_So%s" % (c, n)], rtol=1e-8, atol=1e-12), n
        assert np.allclose(r["Kla_memory"][j], g["learn_c%d_Kla%s" % (c, n)], rtol=1e-10, atol=1e-12), n


def test_env_chain_first_cycle_equals_learning_chain(g):
    """Both chains start from the module's cycle-0 memories: their first cycle is the same computation."""
    if np.array_equal(g["actions"], g["actions_learn"]):
        assert np.array_equal(g["env_c0_x_last"], g["learn_c0_x_last"])
        assert np.array_equal(g["env_c0_u3"], g["learn_c0_u3"])
    assert not np.array_equal(g["env_c1_u3"], g["learn_c1_u3"])
