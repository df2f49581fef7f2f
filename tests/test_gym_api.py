"""The drop-in boundary: the reference's ten env ids, class names, spaces and tuple shapes
(gym_SBR/__init__.py:3-12, gym_SBR_env2.py:58-193, gym_SBR_oneshot.py:98-113,843-1273)."""
import os
import re

import numpy as np
import pytest

import gym_sbr2_b200 as sbr
from gym_sbr2_b200 import _abi

REF_IDS = {"SBR-v0": "SbrEnv", "SBR-v1": "SbrEnv1", "SBR-v2": "SbrEnv2", "SBR-v4": "SbrEnv4",
           "SBRCnt-v0": "SbrCnt0", "SBRCnt-v1": "SbrCnt1", "SBRCnt-v2": "SbrCnt2", "SBRCntMA-v1": "SbrCntMA1",
           "SBROS-v1": "SbrOS", "SBROS-v2": "SbrOS1"}


def test_all_ten_reference_ids_registered_under_reference_class_names():
    assert set(sbr.spec_ids()) == set(REF_IDS)
    import gym_sbr2_b200.envs as envs
    for env_id, cls in REF_IDS.items():
        assert sbr.registry[env_id]["entry_point"] == "gym_sbr2_b200.envs:" + cls
        assert hasattr(envs, cls)


SERVED = tuple(REF_IDS)


def test_every_reference_id_is_served_and_disclosures_are_recorded():
    """All ten ids construct a CUDA-backed env; the table keeps, per id, whether the reference's own step() can run."""
    assert all(row[2] for row in sbr.ENV_TABLE.values())
    src = open(os.path.join(os.path.dirname(sbr.__file__), "registration.py")).read()
    for env_id in ("SBR-v0", "SBR-v1", "SBR-v4", "SBRCnt-v0"):
        assert re.search(r"supported with a disclosure[^\n]*\n(?:\s*#[^\n]*\n)*\s*\"%s\"" % re.escape(env_id), src), env_id


def test_supported_ids_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    for env_id in SERVED:
        with pytest.raises(_abi.SbrLibraryError):
            sbr.make(env_id)


def test_available_actions_mask_matches_reference_rule():
    """gym_SBR_oneshot.py:440-459: +-0.1 on the DO set-point within [0,8], +-5 on the NO3 set-point within [0,15]."""
    from gym_sbr2_b200.envs.single import SbrOS
    masks = SbrOS.get_available_actions(None, [0.05, 12.0], 2, 3)
    assert [list(m) for m in masks] == [[0, 1, 1], [1, 1, 0]]


def test_product_never_imports_oracle_or_scipy():
    pkg = os.path.dirname(sbr.__file__)
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(root, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|scipy)", src, flags=re.M), fn


@pytest.mark.gpu
def test_sbr_v2_known_answer_through_make(built, cuda_device, golden_v2):
    """np.random.seed(0); reset(); step([.25,.25,.25]) -> reward 3.0758784413909894."""
    env = sbr.make("SBR-v2")
    assert env.action_space.shape == (3,) and env.observation_space.shape == (3,)
    np.random.seed(0)
    obs0 = env.reset()
    assert isinstance(obs0, np.ndarray) and obs0.shape == (3,)
    assert np.allclose(obs0, [1.27614847334958, 1.8235276185867406, 1.1067801724189357], rtol=1e-13)
    assert np.array_equal(env.influent_mixed, golden_v2["influent"][0])           # same RNG consumption
    obs, reward, done, info = env.step([0.25, 0.25, 0.25])
    assert done is True and info == {} and isinstance(reward, float)
    assert abs(reward - 3.0758784413909894) <= 1e-5 * 3.08
    assert np.allclose(obs, [0.66, 262.1790959812511, 0.02970193763645346], rtol=1e-5)
    # like the reference, a second step replays the cycle from x0_init (gym_SBR_env2.py:88-99)
    obs2, reward2, _, _ = env.step([0.25, 0.25, 0.25])
    assert reward2 == reward and np.array_equal(obs, obs2)


@pytest.mark.gpu
def test_sbros_v1_episode_through_make(built, cuda_device):
    from test_oracle_golden_os import load_episode
    g = load_episode("seed0_const")
    env = sbr.make("SBROS-v1")
    np.random.seed(0)
    obs = env.reset()
    assert isinstance(obs, tuple) and len(obs) == 2 and len(obs[0]) == 9 and len(obs[1]) == 9
    assert np.array_equal(env.influent_mixed[1:], g["influent"][1:])    # [0] is overwritten with the fill flow (:287)
    assert np.allclose(obs[0], g["reset_obs_do"], rtol=1e-5, atol=1e-7)
    total, k = 0.0, 0
    while True:
        out = env.step([2.0, 5.0])
        assert len(out) == 5                                   # (obs, state, reward, done, info)
        obs, state, reward, done, info = out
        assert len(state) == 15 and isinstance(done, bool) and info == {}
        assert abs(reward - g["reward"][k]) <= 1e-5 * abs(g["reward"][k]) + 1e-9, k
        total += reward
        k += 1
        if done:
            break
    assert k == 463 and abs(total - (-0.878967)) < 1e-5


@pytest.mark.gpu
def test_sbr_v4_episode_through_make(built, cuda_device):
    """SBR-v4: standard 4-tuple, 14-dim obs = x / x_1, 493 steps per episode; same RNG consumption as the reference
    (np.random.choice(8, 1) then the scenario's draws).  Reference run with numpy < 1.18 linspace semantics."""
    from test_oracle_golden_v4 import load_v4
    from test_twin_parity_v4 import V4_SO_SLACK, v4_state_close
    g = load_v4("seed2_walk")
    env = sbr.make("SBR-v4")
    assert env.action_space.shape == (1,) and env.observation_space.shape == (14,)
    np.random.seed(int(g["seed"]))
    obs0 = env.reset()
    assert obs0.shape == (1, 14) and np.allclose(obs0[0], g["reset_obs"], rtol=1e-13)
    k = 0
    while True:
        out = env.step(g["action"][k])
        assert len(out) == 4
        obs, reward, done, info = out
        assert obs.shape == (14,) and isinstance(done, bool) and info == {}
        ok, worst = v4_state_close(obs, g["state"][k], so_slack=V4_SO_SLACK)
        assert ok, (k, worst)
        assert abs(reward - g["reward"][k]) <= 1e-5 * abs(g["reward"][k]) + 1e-9, k
        k += 1
        if done:
            break
    assert k == 493


@pytest.mark.gpu
@pytest.mark.parametrize("env_id,kind,name", [("SBRCnt-v0", "cnt0", "seed1_up"), ("SBRCnt-v1", "cnt1", "seed1_up"),
                                             ("SBRCnt-v2", "cnt2", "seed1_up"), ("SBRCntMA-v1", "ma1", "seed0_up"),
                                             ("SBROS-v2", "os2", "seed1_walk")])
def test_cnt_family_episode_through_make(built, cuda_device, env_id, kind, name):
    """The five ids whose reference step() dies in its reward module: tuple shapes of the reference, the same RNG
    consumption (np.random.seed before reset), observations against the reference run with the repaired reward."""
    from test_twin_parity_cnt import load_cnt, obs_close
    g = load_cnt(kind, name)
    env = sbr.make(env_id)
    np.random.seed(int(g["seed"]))
    obs0 = env.reset()
    assert np.array_equal(env.influent_mixed[1:], g["influent"][1:])
    if kind == "os2":
        assert isinstance(obs0, tuple) and len(obs0[0]) == 9 and len(obs0[1]) == 9
        flat0 = np.concatenate(obs0)
    else:
        assert obs0.shape == ((1, 7) if kind == "cnt0" else (5,))
        flat0 = obs0.reshape(-1)
    assert obs_close(flat0, g["reset_obs"], kind)[0]
    k = 0
    while True:
        out = env.step(g["action"][k] if kind == "os2" else g["action"][k][:1])
        if kind == "os2":
            assert len(out) == 5
            obs, state, reward, done, info = out
            assert state.shape == (15,)
            flat = np.concatenate(obs)
        else:
            assert len(out) == 4
            obs, reward, done, info = out
            flat = obs.reshape(-1)
        assert isinstance(done, bool) and done == bool(g["done"][k]) and info == {}
        ok, worst = obs_close(flat, g["obs"][k], kind)
        assert ok, (k, worst)
        assert reward == g["reward"][k], k
        k += 1
        if done:
            break
    assert k == int(g["n_steps"])
    # trajectory(): one record after the fill phase and one per step(), at the states the next step continues from
    tr = env.trajectory()
    assert len(tr["t_t"]) == k + 1 and tr["x_t"].shape == (k + 1, 14) and len(tr["state_t"]) >= k
    from gym_sbr2_b200 import parity
    assert parity.state_close(tr["x_t"][0], g["x_fill"], atol_frac=1e-7)[0]
    assert parity.state_close(tr["x_t"][k // 2], g["x_cont"][k // 2 - 1], atol_frac=1e-7)[0]
    assert abs(tr["u_DO_t"][k // 2] - g["u_do"][k // 2 - 1]) < 1e-12


@pytest.mark.gpu
def test_sbr_v0_batch_to_batch_env_through_make(built, cuda_device):
    """`SBR-v0`: reset() runs cycle 0 once, step() = batch-to-batch update + one feed-forward cycle; the reference's tuple
    shapes (gym_SBR_env0.py:150-236).  Numbers are pinned in tests/test_gpu_ilc.py."""
    env = sbr.make("SBR-v0")
    assert env.action_space.shape == (3,) and env.observation_space.shape == (14,)
    np.random.seed(0)
    obs0 = env.reset()
    assert isinstance(obs0, np.ndarray) and obs0.shape == (14,) and obs0[0] == 1.0
    assert np.array_equal(env.reset(), obs0)                  # the reference's reset() does not touch the plant
    obs, reward, done, info = env.step([2.0, 2.5, 1.5])
    assert obs.shape == (14,) and done is True and info == {} and np.isfinite(reward)
    u1 = env.info["u_batch"].copy()
    assert np.abs(u1).max() > 0 and int(env.info["status"]) == 0
    obs2, reward2, _, _ = env.step([7.0, -1.0, 1.5])          # clipped to [0, 5]
    assert np.isfinite(obs2).all() and not np.array_equal(env.info["u_batch"], u1)


@pytest.mark.gpu
def test_sbr_v1_env_through_make(built, cuda_device):
    """`SBR-v1`: reset() returns the observation of the module's initial state and never moves the plant; step() = one
    feedback-PID cycle from the carried-over state (gym_SBR_env1.py:105-175).  Numbers are pinned in tests/test_gpu_ilc.py."""
    env = sbr.make("SBR-v1")
    assert env.action_space.shape == (3,) and env.observation_space.shape == (14,)
    np.random.seed(0)
    obs0 = env.reset()
    assert obs0.shape == (14,) and obs0[0] == 1.0
    obs, reward, done, info = env.step([2.0, 2.0, 2.0])
    assert obs.shape == (14,) and done is True and info == {} and np.isfinite(reward)
    x1 = env.info["x_last"].copy()
    assert np.array_equal(env.reset(), obs)                   # reset() does not touch the plant
    env.step([2.0, 2.0, 2.0])
    assert not np.array_equal(env.info["x_last"], x1) and int(env.info["status"]) == 0


@pytest.mark.gpu
def test_sbr_v2_trajectory_through_make(built, cuda_device):
    """SbrEnv2.trajectory(): (t, x, kla) of the last step's cycle in the shapes of SBR_model_FB.run's `t`, `x`."""
    env = sbr.make("SBR-v2")
    np.random.seed(0)
    env.reset()
    with pytest.raises(RuntimeError):
        env.trajectory()
    obs, reward, done, info = env.step([0.25, 0.5, 0.75])
    t, x, kla = env.trajectory()
    assert isinstance(t, list) and len(t) == 529 and x.shape == (14, 529) and kla.shape == (529,)
    assert t[0] > 0 and abs(t[-1] - 0.5) < 1e-2 and all(b >= a for a, b in zip(t, t[1:]))
    assert np.allclose(x[:, -1], env.info["x_last"], rtol=1e-9)
    assert abs(kla[72:295].mean() - env.info["kla3_mean"]) < 1e-9 * max(1.0, abs(env.info["kla3_mean"]))
