"""Pin the SBR-v4 oracle (oracle/sbr_oracle.py::SbrEnv4Oracle) to whole episodes of the reference `SbrEnv4`
(gym_SBR_env4.py) recorded by oracle/make_golden_v4.py.

DISCLOSURE: the reference's step() raises TypeError on numpy >= 1.18 (float `num` in np.linspace, gym_SBR_env4.py:286,
921,982,1207).  The fixtures come from the UNMODIFIED source run with numpy < 1.18 linspace semantics restored for that
module (num -> int(num)); see the header of oracle/make_golden_v4.py."""
import glob
import os

import numpy as np
import pytest

from gym_sbr2_b200 import influent
from oracle import sbr_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
V4_EPISODES = sorted(os.path.basename(f)[len("sbr_v4_"):-4] for f in glob.glob(os.path.join(GOLDEN, "sbr_v4_*.npz")))


def load_v4(name):
    return np.load(os.path.join(GOLDEN, "sbr_v4_%s.npz" % name), allow_pickle=True)


def test_fixture_inventory():
    assert len(V4_EPISODES) == 6
    g = load_v4("seed0_zero")
    assert int(g["n_steps"]) == 493 and list(np.nonzero(g["done"])[0]) == [492]
    bt = list(g["batch_type"])
    assert bt == [0] * 26 + [1] * 466 + [2]               # fill while t < 0.021, react up to t_memory5[-1], then the rest
    assert int(g["n_kla"]) == 494 and int(g["odeint_warnings"]) == 0


@pytest.mark.parametrize("name", V4_EPISODES)
def test_oracle_reproduces_reference_episode(name):
    """Same LSODA calls on the same grids: the oracle reproduces the reference to the last bit on this toolchain."""
    g = load_v4(name)
    o = O.SbrEnv4Oracle()
    s0 = o.reset(np.concatenate([[0.66], g["influent"][1:]]))
    assert np.array_equal(s0, g["reset_obs"])
    assert o.infl[0] == g["influent"][0]
    for k in range(int(g["n_steps"])):
        st, r, done = o.step(float(g["action"][k]))
        assert done == bool(g["done"][k]) and o.batch_type == int(g["batch_type"][k]), k
        assert np.allclose(st, g["state"][k], rtol=1e-9, atol=1e-12), k
        assert abs(r - g["reward"][k]) <= 1e-9 * abs(g["reward"][k]) + 1e-12, k
        assert abs(o.u - g["u"][k]) < 1e-15 and abs(o.Kla[-1] - g["kla"][k]) <= 1e-8 * max(1.0, abs(g["kla"][k])), k
        if not done:
            assert o.t == g["t"][k], k
    assert np.isclose(o.Qw, float(g["Qw"]), rtol=1e-9) and len(o.Kla) == int(g["n_kla"])
    assert np.isclose(sum(o.Kla), float(g["kla_sum"]), rtol=1e-9)
    assert o.idle_pts == 361 or o.idle_pts > 300


@pytest.mark.parametrize("name", V4_EPISODES)
def test_random_scenario_draw_consumes_rng_like_the_reference(name):
    """reset(): buffer_tank(np.random.choice(8, 1)) (gym_SBR_env4.py:104) -- same RNG consumption, same influent."""
    g = load_v4(name)
    np.random.seed(int(g["seed"]))
    sw, infl = influent.sample_numpy_random_scenario()
    assert 0 <= sw < 8 and np.array_equal(infl[1:], g["influent"][1:]) and infl[0] == 0.66
