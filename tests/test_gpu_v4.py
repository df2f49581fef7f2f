"""GPU parity tests of the SBR-v4 path (sbr_v4_reset / sbr_v4_step through the C ABI): whole episodes of the
reference `SbrEnv4` (numpy < 1.18 linspace semantics, see test_oracle_golden_v4.py), every env of a 4096 batch against
the g++ twin, and the vector-env API."""
import numpy as np
import pytest
import torch

from gym_sbr2_b200 import _abi, core, schedule
from gym_sbr2_b200.vec_env import SbrV4VecEnv
from oracle.twin import binding as twin
from test_oracle_golden_v4 import V4_EPISODES, load_v4
from test_twin_parity_v4 import check_v4, run_v4

pytestmark = pytest.mark.gpu


class GpuV4Batch(object):
    """numpy-in / numpy-out adapter over the CUDA entry points with the interface of oracle.twin.binding.V4Batch."""

    def __init__(self, n, device, mode=_abi.MODE_DP45):
        self.n, self.device, self.mode = n, device, mode
        self.params, self.sched, self.tol = _abi.default_params(), schedule.os_schedule(), _abi.make_tol()
        self.buf = core.V4Buffers(n, device)

    def _dev(self, a, dtype=torch.float64):
        return None if a is None else core.soa1(torch.as_tensor(np.ascontiguousarray(a)).to(self.device, dtype))

    def _sync(self):
        torch.cuda.synchronize()
        b = self.buf
        self.st, self.status = b.st.cpu().numpy(), b.status.cpu().numpy()
        self.counters, self.done = b.counters.cpu().numpy().astype(np.uint32), b.done.cpu().numpy()

    def reset(self, influent, x0=None, mask=None):
        self.influent = self._dev(influent)
        core.v4_reset(self.buf, self.influent, self.params, x0=self._dev(x0), mask=self._dev(mask, torch.uint8))
        self._sync()
        return self.buf.obs.cpu().numpy()

    def step(self, action):
        a = torch.as_tensor(np.ascontiguousarray(action, dtype=np.float64).reshape(-1)).to(self.device)
        b = core.v4_step(self.buf, self.influent, a, self.params, self.sched, mode=self.mode, tol=self.tol)
        self._sync()
        return b.obs.cpu().numpy(), b.reward.cpu().numpy(), self.done.copy()


def test_dp45_episodes_match_reference(built, cuda_device):
    G, b, ob0, rec = run_v4(lambda n: GpuV4Batch(n, cuda_device), V4_EPISODES)
    check_v4(G, b, ob0, rec, V4_EPISODES)


@pytest.mark.parametrize("mode,tol", [(_abi.MODE_RK4, 1e-9), (_abi.MODE_DP45, 2e-6)])
def test_4096_envs_every_env_against_cpu_twin(built, cuda_device, mode, tol):
    """4096 envs, per-env scenario and influent, random delta actions, 60 steps across the fill -> react switch."""
    from gym_sbr2_b200 import influent
    n = 4096
    rng = np.random.RandomState(31)
    infl = np.stack([influent.mix_numpy(rng.randint(8), rng.randn(48)) for _ in range(128)], axis=1)
    infl = np.tile(infl, (1, n // 128)).copy()
    infl[0] = schedule.os_fill_flow(_abi.default_params().Qin)
    g, c = GpuV4Batch(n, cuda_device, mode=mode), twin.V4Batch(n, mode=mode)
    og, oc = g.reset(infl), c.reset(infl)
    assert np.array_equal(og, oc) or np.allclose(og, oc, rtol=1e-15, atol=0)
    for k in range(60):
        act = rng.uniform(-1, 1, n) * (0.3 if k > 5 else 1.0)
        (sg, rg, dg), (sc, rc, dc) = g.step(act), c.step(act)
        assert np.array_equal(dg, dc) and np.array_equal(g.status, c.status)
        assert np.all(np.abs(sg - sc) <= tol * np.abs(sc) + tol * 1e-2), (k, np.abs(sg - sc).max())
        assert np.allclose(rg, rc, rtol=max(tol, 1e-9) * 10, atol=1e-10), k
        if mode == _abi.MODE_RK4:
            assert np.array_equal(g.counters, c.counters)


def test_vec_env_api_scenarios_autoreset(built, cuda_device):
    n = 64
    env = SbrV4VecEnv(n, device=cuda_device, seed=7)
    obs = env.reset()
    assert obs.shape == (n, 14) and bool(torch.isfinite(obs).all())
    assert int(env.scenario.min()) >= 0 and int(env.scenario.max()) <= 7 and len(torch.unique(env.scenario)) > 3
    assert float(obs[:, 0].min()) == 1.0 == float(obs[:, 0].max())          # (Qin + IV) / x_1[0] = 1.32 / 1.32
    a = torch.full((n, 1), 0.05, dtype=torch.float64, device=cuda_device)
    for k in range(493):
        obs, reward, done, info = env.step(a if k < 60 else torch.zeros_like(a))
        assert bool(done.all()) == (k == 492)
    assert float(info["episode_steps"].min()) == 493 and float(info["u"].max()) <= 8.0
    assert bool(torch.isfinite(info["Qw"]).all()) and int(info["status"].max()) == 0
    obs2, reward, done, info = env.step(a)                                       # finished: no-op
    assert torch.equal(obs2, obs) and float(reward.abs().max()) == 0.0
    env.autoreset = True
    obs3, reward, done, info = env.step(a)
    assert not bool(done.any()) and float(info["episode_steps"].max()) == 1 and bool(info["restarted"].all())


def test_state_placement_does_not_change_results(built, cuda_device):
    """order='steps': the persistent state lives in slots re-sorted by the previous step's RHS count (first at step 8,
    then every 8 / 32 steps), everything the caller sees stays indexed by env.  A whole episode with per-env scenarios
    and random delta actions, a masked reset in the middle: observations, rewards, done flags, info and the final
    state are bit-identical to the identity placement."""
    n = 4096
    a = SbrV4VecEnv(n, device=cuda_device, seed=11, order="steps")
    b = SbrV4VecEnv(n, device=cuda_device, seed=11, order="none")
    assert a.order == "steps" and SbrV4VecEnv(n, device=cuda_device).order == "steps" and b.order == "none"
    oa, ob = a.reset(), b.reset()
    assert torch.equal(oa, ob)
    g = torch.Generator(device=cuda_device).manual_seed(2)
    mask = (torch.arange(n, device=cuda_device) % 5 == 0)
    for k in range(493):
        act = 0.2 * torch.randn(n, dtype=torch.float64, device=cuda_device, generator=g) + 0.02
        if k == 100:
            a.reset(mask=mask); b.reset(mask=mask)
            assert a._slot_env is not None                      # the masked reset ran with the placement in force
        ra, rb = a.step(act), b.step(act)
        assert torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1]) and torch.equal(ra[2], rb[2]), k
        if k % 50 == 0 or k > 485:
            for key in ("t", "u", "episode_return", "episode_steps", "counters", "status"):
                assert torch.equal(ra[3][key], rb[3][key]), (k, key)
    assert a._slot_env is not None and not torch.equal(a._slot_env.long(), torch.arange(n, device=cuda_device))
    assert bool(ra[2][~mask].all()) and not bool(ra[2][mask].any())          # the restarted envs are 101 steps behind
    qa = ra[3]["Qw"]
    assert torch.equal(torch.nan_to_num(qa, nan=-1.0), torch.nan_to_num(rb[3]["Qw"], nan=-1.0))
    a.unsort()
    assert a._slot_env is None
    assert torch.equal(a.buf.st.view(torch.int64), b.buf.st.view(torch.int64))


@pytest.mark.parametrize("K,resort", [(8, 2), (5, 1), (8, 0)])
def test_fused_rollout_matches_stepwise_rollout(built, cuda_device, K, resort):
    """sbr_v4_rollout_k (K steps per launch, 14 -> 32 -> 1 policy head in-kernel, every buffer in slot order and fully
    re-sorted by RHS count between launches) against the step-by-step rollout [sbr_policy_mlp, sbr_v4_step]: per-env
    episode returns and final states agree to rounding (two instantiations of the same arithmetic), done everywhere."""
    from gym_sbr2_b200 import rollout
    n = 1000
    policy = rollout.TinyPolicy(cuda_device, n_in=14, lo=(-0.1,), span=(0.4,), seed=3)
    env_s = SbrV4VecEnv(n, device=cuda_device, seed=21)
    step = rollout.collect_episode_v4(env_s, policy)
    env_f = SbrV4VecEnv(n, device=cuda_device, seed=21)
    fused = rollout.collect_episode_v4_fused(env_f, policy, K=K, resort_every=resort)
    assert bool(step["all_done"]) and bool(fused["all_done"]) and fused["steps"] == 493
    assert int(fused["status"].max()) == 0
    env_s.unsort()
    rs, rf = step["returns"], fused["returns"]
    assert float((rs - rf).abs().max()) <= 1e-9 * float(rs.abs().max())
    assert torch.allclose(env_f.buf.st[:14], env_s.buf.st[:14], rtol=1e-9, atol=1e-12)
    assert torch.equal(env_f.buf.st[_abi.V4_STEPS], env_s.buf.st[_abi.V4_STEPS])
    # the policy moves the set-point (state feedback through 14 observations) and it stays inside [0, 8]
    u = env_f.buf.st[_abi.V4_U]
    assert float(u.min()) >= 0.0 and float(u.max()) <= 8.0 and float(u.max()) > 0.5
    assert torch.equal(env_f.buf.st[_abi.V4_U], env_s.buf.st[_abi.V4_U])


@pytest.mark.parametrize("resort", [4, 32])
def test_slot_ordered_stepwise_rollout_is_bit_identical(built, cuda_device, resort):
    """collect_episode_v4_sorted: one policy launch and one sbr_v4_step per env.step with every buffer in slot order and a
    full re-sort every few steps -- the same kernel on the same per-env inputs in another order: identical bits, in env
    order again at the end."""
    from gym_sbr2_b200 import rollout
    n = 777
    policy = rollout.TinyPolicy(cuda_device, n_in=14, lo=(-0.1,), span=(0.4,), seed=3)
    env_s = SbrV4VecEnv(n, device=cuda_device, seed=21)
    step = rollout.collect_episode_v4(env_s, policy)
    env_s.unsort()
    env_o = SbrV4VecEnv(n, device=cuda_device, seed=21)
    srt = rollout.collect_episode_v4_sorted(env_o, policy, resort_every=resort)
    assert bool(srt["all_done"]) and srt["steps"] == 493 and int(srt["status"].max()) == 0
    assert torch.equal(step["returns"], srt["returns"])
    assert torch.equal(env_o.buf.st, env_s.buf.st) and torch.equal(env_o.buf.obs, env_s.buf.obs)
