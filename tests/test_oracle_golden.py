"""Pin the oracle (oracle/sbr_oracle.py) to outputs of the UNMODIFIED reference (tests/golden/, produced by
oracle/make_golden.py).  The oracle calls the same scipy LSODA with the same grids, so agreement is expected
at rounding level; the asserted bound (1e-9) is four orders tighter than the parity tolerance it is used for."""
import numpy as np
import pytest

from oracle import sbr_oracle as O

PIN_RTOL = 1e-9


def test_versions_recorded(golden_v2):
    v = [str(s) for s in golden_v2["versions"]]
    assert len(v) == 3 and all(v)


def test_rhs_three_tails_match_reference(stage_samples):
    s = stage_samples
    ec_conc = float(s["ec_conc"])
    for i in range(len(s["x"])):
        x, kla, ec = s["x"][i], float(s["kla"][i]), float(s["ec"][i])
        for mine, ref in ((O.rhs_react(x, 0.0, kla), s["d_react"][i]),
                          (O.rhs_fill(x, 0.0, kla, s["load"]), s["d_fill"][i]),
                          (O.rhs_react_ec(x, 0.0, kla, ec, ec_conc), s["d_ec"][i])):
            assert np.allclose(mine, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())


def test_do_saturation_constant(stage_samples):
    assert O.SO_SAT == float(stage_samples["so_sat"]) == 8.000000000006622


def test_settle_and_draw_match_reference(stage_samples):
    s = stage_samples
    t0 = 0.4169166666666667
    for i in range(len(s["settle_x"])):
        sX, Xf = O.settle(s["settle_x"][i], t0, t0 + 0.5 * 0.083)
        assert np.allclose(sX, s["settle_sX"][i], rtol=PIN_RTOL, atol=0)
        assert Xf == s["settle_Xf"][i]
        x7, Qw, EQI, eff, status = O.draw(s["settle_x"][i], s["settle_sX"][i], s["settle_Xf"][i])
        assert status == 0
        assert np.allclose(x7, s["draw_x7"][i], rtol=1e-13, atol=0)
        assert np.isclose(Qw, s["draw_Qw"][i], rtol=1e-12)
        assert np.isclose(EQI, s["draw_EQI"][i], rtol=1e-13)
        assert np.allclose(eff, s["draw_eff"][i], rtol=1e-13)


def test_known_answer_vector_seed0(golden_v2):
    """SURVEY.md 8c known-answer: seed 0, action [.25,.25,.25] -> reward 3.0758784413909894."""
    g = golden_v2
    i = 0
    assert int(g["seed"][i]) == 0 and np.allclose(g["action"][i], 0.25)
    assert g["reward"][i] == 3.0758784413909894
    assert np.allclose(g["reset_obs"][i], [1.27614847334958, 1.8235276185867406, 1.1067801724189357], rtol=0, atol=0)
    out = O.sbr_v2_step(g["action"][i], g["influent"][i])
    assert abs(out["reward"] - 3.0758784413909894) <= PIN_RTOL * 3.08
    assert out["n_intervals"] == [24, 48, 223, 186, 11, 0, 0, 36]


@pytest.mark.parametrize("chunk", range(4))
def test_whole_cycle_matches_reference(golden_v2, chunk):
    g = golden_v2
    idx = list(range(len(g["seed"])))[chunk::4]
    for i in idx:
        out = O.sbr_v2_step(g["action"][i], g["influent"][i])
        assert np.allclose(out["x_last"], g["x_last"][i], rtol=PIN_RTOL, atol=1e-14), i
        assert np.isclose(out["reward"], g["reward"][i], rtol=PIN_RTOL), i
        assert np.allclose(out["obs"], g["obs"][i], rtol=PIN_RTOL), i
        assert np.isclose(out["Qw"], g["Qw"][i], rtol=PIN_RTOL), i
        assert np.isclose(out["EQI"], g["EQI"][i], rtol=PIN_RTOL), i
        assert np.allclose(out["eff"], g["eff"][i], rtol=PIN_RTOL), i
        assert out["done"] is True and bool(g["done"][i]) is True
        assert (len(out["kla"][2]), len(out["kla"][4]), len(out["kla"][7])) == (g["n3"][i], g["n5"][i], g["n8"][i])
        assert np.isclose(np.mean(out["kla"][2]), g["kla3_mean"][i], rtol=PIN_RTOL, atol=1e-12), i
        assert np.isclose(np.mean(out["kla"][7]), g["kla8_mean"][i], rtol=PIN_RTOL, atol=1e-12), i
        assert np.allclose(O.sbr_v2_reset_obs(g["influent"][i]), g["reset_obs"][i], rtol=1e-15, atol=0), i
        # the fill flow the reference writes into influent_mixed[0] at step time (gym_SBR_env2.py:144)
        assert g["influent_step"][i][0] == O.fill_flow()
