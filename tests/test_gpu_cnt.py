"""GPU parity tests of `SBRCnt-v0/1/2`, `SBRCntMA-v1`, `SBROS-v2` (sbr_cnt_reset / sbr_cnt_step through the C ABI):
whole episodes of the unmodified reference env modules (reward repaired as oracle/make_golden_cnt.py discloses), every
env of a 2048 batch against the g++ twin, and the vector-env API."""
import numpy as np
import pytest
import torch

from gym_sbr2_b200 import _abi, cnt, schedule
from oracle.twin import binding as twin
from test_twin_parity_cnt import CNT_EPISODES, run_cnt

pytestmark = pytest.mark.gpu


class GpuCntBatch(object):
    """numpy-in / numpy-out adapter over the CUDA entry points with the interface of oracle.twin.binding.CntBatch."""

    def __init__(self, kind, n, device, mode=_abi.MODE_DP45):
        self.n, self.device, self.mode = n, device, mode
        self.cfg = cnt.cnt_config(kind)
        self.params, self.sched, self.tol = _abi.default_params(), schedule.os_schedule(), _abi.make_tol()
        self.buf = cnt.CntBuffers(n, device, cnt.OBS_ROWS[kind])

    def _sync(self):
        torch.cuda.synchronize()
        b = self.buf
        self.st, self.status = b.st.cpu().numpy(), b.status.cpu().numpy()
        self.counters, self.done = b.counters.cpu().numpy().astype(np.uint32), b.done.cpu().numpy()

    def reset(self, influent, x0=None, mask=None):
        infl = torch.as_tensor(np.ascontiguousarray(influent)).to(self.device, torch.float64)
        if self.n == 1:
            infl = infl.reshape(14, 1).contiguous()
        cnt.cnt_reset(self.cfg, self.buf, infl, self.params, self.sched, mode=self.mode, tol=self.tol)
        self._sync()
        return self.buf.obs.cpu().numpy()

    def step(self, action):
        a = torch.as_tensor(np.ascontiguousarray(action, dtype=np.float64)).to(self.device)
        b = cnt.cnt_step(self.cfg, self.buf, a, self.params, self.sched, mode=self.mode, tol=self.tol)
        self._sync()
        return b.obs.cpu().numpy(), b.reward.cpu().numpy(), self.done.copy()


@pytest.mark.parametrize("episode", CNT_EPISODES)
def test_dp45_episodes_match_reference(built, cuda_device, episode):
    run_cnt(lambda kind: GpuCntBatch(kind, 1, cuda_device), episode)


@pytest.mark.parametrize("kind", sorted(cnt.KINDS))
def test_2048_envs_every_env_against_cpu_twin(built, cuda_device, kind):
    """2048 envs with per-env influent and random actions over the start of the episode (the whole-phase solve of
    cnt1 / cnt2, the anoxic -> aerobic switch of ma1 / os2): every env against the g++ build of the same arithmetic."""
    from gym_sbr2_b200 import influent
    n = 2048
    rng = np.random.RandomState(41)
    infl = np.stack([influent.mix_numpy(0, rng.randn(48)) for _ in range(64)], axis=1)
    infl = np.tile(infl, (1, n // 64)).copy()
    infl[0] = schedule.os_fill_flow(_abi.default_params().Qin)
    g = GpuCntBatch(kind, n, cuda_device)
    c = twin.CntBatch(cnt.cnt_config(kind), n, cnt.OBS_ROWS[kind], mode=_abi.MODE_DP45)
    og, oc = g.reset(infl), c.reset(infl)
    assert np.allclose(og, oc, rtol=1e-9, atol=1e-12)
    steps = 70 if kind in ("ma1", "os2", "cnt0") else 12
    for k in range(steps):
        act = np.zeros((2, n))
        if kind == "os2":
            act[0], act[1] = rng.uniform(0.5, 4.0, n), rng.uniform(0.0, 0.4, n)
        else:
            act[0] = rng.uniform(-1, 1, n) * (0.05 if kind == "cnt0" else 0.5)
        (sg, rg, dg), (sc, rc, dc) = g.step(act), c.step(act)
        assert np.array_equal(dg, dc) and np.array_equal(g.status, c.status), k
        assert np.all(np.abs(g.st[:14] - c.st[:14]) <= 2e-6 * np.abs(c.st[:14]) + 2e-8), (k, np.abs(g.st[:14] - c.st[:14]).max())
        assert np.all(np.abs(sg - sc) <= 2e-6 * np.abs(sc) + 2e-6), k
        assert np.mean(rg == rc) > 0.999, k                     # a So within 1e-6 of a reward threshold may flip its bin


@pytest.mark.parametrize("kind", sorted(cnt.KINDS))
def test_vec_env_episode_api(built, cuda_device, kind):
    n = 96
    env = cnt.SbrCntVecEnv(kind, n, device=cuda_device, seed=5)
    obs = env.reset()
    if kind == "os2":
        assert obs[0].shape == (n, 9) and obs[1].shape == (n, 9)
    else:
        assert obs.shape == (n, cnt.OBS_ROWS[kind]) and bool(torch.isfinite(obs).all())
    g = torch.Generator(device=cuda_device).manual_seed(3)
    for k in range(env.max_episode_steps):
        if kind == "os2":
            a = torch.stack([1.0 + 2.0 * torch.rand(n, dtype=torch.float64, device=cuda_device, generator=g),
                             torch.zeros(n, dtype=torch.float64, device=cuda_device)], dim=1)
        else:
            # raise the DO set-point to 2 g/m3 over the first 8 aerobic steps, hold it otherwise
            up = range(60, 68) if kind == "ma1" else range(1, 9)
            a = torch.full((n, 1), (0.05 if kind == "cnt0" else 0.25) if k in up else 0.0, dtype=torch.float64,
                           device=cuda_device)
            if kind in ("ma1", "cnt2") and k == 0:
                a.fill_(-2.0)                    # carbon set-point to 0 before the controller can run away
        out = env.step(a)
        done, info = out[-2], out[-1]
        assert bool(done.all()) == (k == env.max_episode_steps - 1), k
    assert int(info["status"].max()) == 0 and bool(torch.isfinite(info["Qw"]).all())
    assert float(info["episode_steps"].min()) == env.max_episode_steps
    assert float(env.buf.st[0].max()) < 1.4                    # the reactor stayed physical
    before = env.buf.st.clone()
    out = env.step(a)                                           # finished: no-op
    assert torch.equal(torch.nan_to_num(env.buf.st), torch.nan_to_num(before)) and float(out[-3].abs().max()) == 0.0
    sd = env.state_dict()
    env2 = cnt.SbrCntVecEnv(kind, n, device=cuda_device, seed=5)
    env2.load_state_dict(sd)
    assert torch.equal(torch.nan_to_num(env2.buf.st), torch.nan_to_num(env.buf.st))
    assert torch.equal(env2.buf.obs, env.buf.obs)


@pytest.mark.parametrize("kind", sorted(cnt.KINDS))
def test_random_actions_against_scipy_oracle(built, cuda_device, kind):
    """A batch of 6 envs with fresh influent draws and random per-env action sequences (not the fixtures) on the GPU
    against the scipy restatement of the reference env (oracle.SbrCntOracle, pinned to the reference's own episodes by
    tests/test_oracle_golden_cnt.py), every env replayed on the CPU."""
    import warnings
    from gym_sbr2_b200 import influent, parity
    from oracle import sbr_oracle as O
    from test_twin_parity_cnt import STATE_ATOL, obs_close, random_plan
    n = 6
    rng = np.random.RandomState(123)
    steps = 130 if kind in ("cnt0", "ma1", "os2") else 50
    infl = np.stack([influent.mix_numpy(0, rng.randn(48)) for _ in range(n)], axis=1)
    plans = [random_plan(kind, rng, steps) for _ in range(n)]
    load = infl.copy()
    load[0] = schedule.os_fill_flow(_abi.default_params().Qin)
    g = GpuCntBatch(kind, n, cuda_device)
    g.reset(load)
    oracles = [O.SbrCntOracle(kind) for _ in range(n)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for j, o in enumerate(oracles):
            o.reset(np.concatenate([[0.66], infl[1:, j]]))
        for k in range(steps):
            act = np.stack([p[k] for p in plans], axis=1)
            ob, r, d = g.step(act)
            for j, o in enumerate(oracles):
                out = o.step(plans[j][k] if kind == "os2" else plans[j][k][:1])
                ok, w = parity.state_close(g.st[:14, j], o.x, atol_frac=STATE_ATOL)
                assert ok, (kind, j, k, w)
                ref_obs = np.concatenate(out[0]) if kind == "os2" else np.asarray(out[0]).reshape(-1)
                assert obs_close(ob[:, j], ref_obs, kind)[0], (kind, j, k)
                assert bool(d[j]) == bool(out[-1])
    assert int(g.status.max()) == 0


@pytest.mark.parametrize("kind", ["cnt1", "os2"])
def test_rk4_mode_against_cpu_twin(built, cuda_device, kind):
    """Fixed-step mode (the reference's output grid as RK4 sub-steps; a whole-phase solve = 9 sub-steps per control interval
    it spans): GPU and g++ builds of the same arithmetic agree to rounding, RHS counters exactly."""
    from gym_sbr2_b200 import influent
    n = 512
    rng = np.random.RandomState(9)
    infl = np.stack([influent.mix_numpy(0, rng.randn(48)) for _ in range(32)], axis=1)
    infl = np.tile(infl, (1, n // 32)).copy()
    infl[0] = schedule.os_fill_flow(_abi.default_params().Qin)
    g = GpuCntBatch(kind, n, cuda_device, mode=_abi.MODE_RK4)
    c = twin.CntBatch(cnt.cnt_config(kind), n, cnt.OBS_ROWS[kind], mode=_abi.MODE_RK4)
    g.reset(infl); c.reset(infl)
    assert np.array_equal(g.counters, c.counters)
    for k in range(6):
        act = np.zeros((2, n))
        act[0] = rng.uniform(0.5, 3.0, n) if kind == "os2" else rng.uniform(0.0, 0.5, n)
        g.step(act); c.step(act)
        assert np.array_equal(g.counters, c.counters), k
        assert np.all(np.abs(g.st[:14] - c.st[:14]) <= 1e-9 * np.abs(c.st[:14]) + 1e-11), k
    # cnt1's first step runs the whole anoxic phase 2 (0.0415 d = 49.8 control intervals -> 50 x 9 sub-steps) plus one
    # control interval of 9 sub-steps
    if kind == "cnt1":
        g2 = GpuCntBatch(kind, 1, cuda_device, mode=_abi.MODE_RK4)
        g2.reset(infl[:, :1].copy())
        g2.step(np.zeros((2, 1)))
        q = cnt.cnt_config(kind)
        n_iv = int(np.ceil((q.tm2_1 - q.tm2_0) / schedule.os_schedule().t_delta))
        assert n_iv == 50 and int(g2.counters[0, 0]) == 4 * (n_iv * 9 + 9)


def test_poisoned_actions_are_flagged_and_masked_reset_restarts_only_the_chosen_envs(built, cuda_device):
    """NaN / huge actions do not crash a launch: the env is flagged (SBR_ST_NONFINITE) or clipped like the reference's
    if / elif chain clips the accumulated set-point; reset(mask) restarts the marked envs and leaves the others alone."""
    n = 64
    env = cnt.SbrCntVecEnv("cnt1", n, device=cuda_device, seed=2)
    env.reset()
    a = torch.zeros((n, 1), dtype=torch.float64, device=cuda_device)
    a[3] = float("nan")
    a[5] = 1e9                                   # clipped to the set-point's upper bound 8
    a[7] = -1e9                                  # ... and to 0
    out = env.step(a)
    info = out[-1]
    assert int(info["status"][3]) & _abi.ST_NONFINITE
    assert float(info["u_do"][5]) == 8.0 and float(info["u_do"][7]) == 0.0
    ok = torch.ones(n, dtype=torch.bool, device=cuda_device)
    ok[3] = False
    assert int(info["status"][ok].max()) == 0
    for _ in range(5):
        env.step(torch.zeros_like(a))
    before = env.buf.st.clone()
    mask = torch.zeros(n, dtype=torch.bool, device=cuda_device)
    mask[3] = mask[10] = True
    env.reset(mask=mask)
    after = env.buf.st
    assert torch.equal(torch.nan_to_num(after[:, ~mask]), torch.nan_to_num(before[:, ~mask]))
    assert float(after[_abi.CNT_STEPS, 3]) == 0.0 and float(after[_abi.CNT_STEPS, 10]) == 0.0
    assert float(after[_abi.CNT_STEPS, 0]) == 6.0 and bool(torch.isfinite(after[:14, 3]).all())
    with pytest.raises(ValueError):
        env.step(torch.zeros((n, 3), dtype=torch.float64, device=cuda_device))
    with pytest.raises(ValueError):
        cnt.SbrCntVecEnv("nope", 4, device=cuda_device)


def cnt_policy(kind, device):
    """Stand-in policy heads on the kind's observation rows: small set-point moves (absolute set-points for os2)."""
    from gym_sbr2_b200 import rollout
    if kind == "os2":
        return rollout.TinyPolicy(device, n_in=18, lo=(0.5, 0.0), span=(3.0, 0.0), seed=4)      # NO3 set-point held at 0
    span = 0.004 if kind == "cnt0" else 0.06
    return rollout.TinyPolicy(device, n_in=cnt.POLICY_INPUTS[kind], lo=(-span / 4,), span=(span,), seed=4)


@pytest.mark.parametrize("kind", sorted(cnt.KINDS))
def test_fused_rollout_matches_stepwise_rollout(built, cuda_device, kind):
    """sbr_cnt_rollout_k (K steps per launch, policy head in-kernel) against [sbr_policy_mlp, sbr_cnt_step] step by step:
    same episode, per-env returns and final state to rounding (two instantiations of one arithmetic)."""
    from gym_sbr2_b200 import rollout
    n = 500
    policy = cnt_policy(kind, cuda_device)
    env_s = cnt.SbrCntVecEnv(kind, n, device=cuda_device, seed=31)
    step = rollout.collect_episode_cnt(env_s, policy)
    env_f = cnt.SbrCntVecEnv(kind, n, device=cuda_device, seed=31)
    fused = rollout.collect_episode_cnt_fused(env_f, policy, K=7)
    assert bool(step["all_done"]) and bool(fused["all_done"])
    assert torch.equal(env_f.buf.st[_abi.CNT_STEPS], env_s.buf.st[_abi.CNT_STEPS])
    assert float(env_f.buf.st[_abi.CNT_STEPS].min()) == env_f.max_episode_steps
    rs, rf = step["returns"], fused["returns"]
    assert float((rs - rf).abs().max()) <= 1e-9 * max(1.0, float(rs.abs().max()))
    assert torch.allclose(env_f.buf.st[:14], env_s.buf.st[:14], rtol=1e-9, atol=1e-12)
    assert torch.allclose(env_f.buf.obs, env_s.buf.obs, rtol=1e-9, atol=1e-12)
    assert torch.allclose(env_f.buf.st[_abi.CNT_U_DO], env_s.buf.st[_abi.CNT_U_DO], rtol=1e-9, atol=1e-12)
    # (the carbon controllers of cnt2 / ma1 run away under a random policy, as in the reference: no physicality claim here)
    assert torch.equal(step["status"], fused["status"])
