"""GPU parity tests of the interval-per-step path (SBROS-v1): kernels sbr_os_reset / sbr_os_step through the C ABI
against whole episodes of the unmodified reference, against the scipy oracle, and (every env of a 4096 batch)
against the g++ twin of the same stepper."""
import numpy as np
import pytest
import torch

from gym_sbr2_b200 import _abi, core, parity, schedule
from gym_sbr2_b200.vec_env import SbrOsVecEnv
from oracle.twin import binding as twin
from test_oracle_golden_os import EPISODES, load_episode
from test_twin_parity_os import check_against_golden, run_episodes

pytestmark = pytest.mark.gpu


class GpuOsBatch(object):
    """numpy-in / numpy-out adapter over the CUDA entry points with the interface of oracle.twin.binding.OsBatch."""

    def __init__(self, n, device, mode=_abi.MODE_DP45, sched=None, tol=None):
        self.n, self.device, self.mode = n, device, mode
        self.params = _abi.default_params()
        self.sched = sched or schedule.os_schedule()
        self.tol = tol or _abi.make_tol()
        self.buf = core.OsBuffers(n, device)

    def _dev(self, a, dtype=torch.float64):
        return None if a is None else core.soa1(torch.as_tensor(np.ascontiguousarray(a)).to(self.device, dtype))

    def _sync(self):
        b = self.buf
        torch.cuda.synchronize()
        self.st = b.st.cpu().numpy()
        self.status = b.status.cpu().numpy()
        self.counters = b.counters.cpu().numpy().astype(np.uint32)
        self.done = b.done.cpu().numpy()

    def reset(self, influent, x0=None, mask=None):
        core.os_reset(self.buf, self._dev(influent), self.params, self.sched, x0=self._dev(x0),
                      mask=self._dev(mask, torch.uint8), mode=self.mode, tol=self.tol)
        self._sync()
        return self.buf.obs_do.cpu().numpy(), self.buf.obs_ec.cpu().numpy()

    def step(self, action):
        b = core.os_step(self.buf, self._dev(action), self.params, self.sched, mode=self.mode, tol=self.tol)
        self._sync()
        return (b.obs_do.cpu().numpy(), b.obs_ec.cpu().numpy(), b.state.cpu().numpy(), b.reward.cpu().numpy(),
                self.done.copy())


def test_dp45_episodes_match_reference(built, cuda_device):
    G, rec = run_episodes(lambda n: GpuOsBatch(n, cuda_device, mode=_abi.MODE_DP45), EPISODES)
    check_against_golden(G, rec, EPISODES)


def test_rk4_episodes_match_reference(built, cuda_device):
    names = ["seed0_const", "seed2_walk", "seed4_random", "seed5_walk"]
    sched = schedule.os_schedule(rk4_sub_interval=20)
    G, rec = run_episodes(lambda n: GpuOsBatch(n, cuda_device, mode=_abi.MODE_RK4, sched=sched), names)
    check_against_golden(G, rec, names)


def test_rk4_reference_grid_counters_and_double_steps(built, cuda_device):
    """The 9/10 output-point pattern of the reference (bit-exact float truncation of the running time) shows in the
    RK4 RHS counters: 32 or 36 per single interval, two intervals at steps 51 and 275."""
    g = load_episode("seed0_const")
    b = GpuOsBatch(1, cuda_device, mode=_abi.MODE_RK4)
    c = twin.OsBatch(1, mode=_abi.MODE_RK4)
    b.reset(g["influent"][:, None]); c.reset(g["influent"][:, None])
    assert b.counters[0, 0] == 4 * 251
    for k in range(463):
        a = g["action"][k][:, None]
        out_g, out_c = b.step(a), c.step(a)
        assert b.counters[0, 0] == c.counters[0, 0], k
        assert out_g[2][0, 0] == out_c[2][0, 0], k                    # t / 0.5: bit-identical running time
        assert bool(out_g[4][0]) == bool(g["done"][k])
    assert b.st[_abi.OS_STEPS, 0] == 463


@pytest.mark.parametrize("mode,tol", [(_abi.MODE_RK4, 1e-9), (_abi.MODE_DP45, 2e-6)])
def test_4096_envs_every_env_against_cpu_twin(built, cuda_device, mode, tol):
    """BASELINE config 2 shape: 4096 envs, per-env influent and random per-step set-points, 60 steps across the
    anoxic -> aerobic switch; every env against the g++ build of the same stepper."""
    from gym_sbr2_b200 import influent
    n = 4096
    rng = np.random.RandomState(21)
    infl = np.stack([influent.mix_numpy(6, rng.randn(48)) for _ in range(128)], axis=1)
    infl = np.tile(infl, (1, n // 128)).copy()
    infl[0] = schedule.os_fill_flow(_abi.default_params().Qin)
    g, c = GpuOsBatch(n, cuda_device, mode=mode), twin.OsBatch(n, mode=mode)
    og, oc = g.reset(infl), c.reset(infl)
    assert np.allclose(og[0], oc[0], rtol=tol, atol=tol) and np.allclose(og[1], oc[1], rtol=tol, atol=tol)
    for k in range(60):
        act = np.stack([8 * rng.rand(n), 15 * rng.rand(n)])
        rg, rc = g.step(act), c.step(act)
        assert np.array_equal(rg[4], rc[4])
        assert np.array_equal(g.status, c.status)
        ok, worst = parity.os_close(rg[2], rc[2], rtol=tol, atol=tol * 1e-2)
        assert ok, (k, worst)
        assert np.allclose(rg[3], rc[3], rtol=max(tol, 1e-9) * 10, atol=1e-10), k
        if mode == _abi.MODE_RK4:
            assert np.array_equal(g.counters, c.counters)


def test_vec_env_api_done_noop_and_autoreset(built, cuda_device):
    g = load_episode("seed0_const")
    n = 8
    env = SbrOsVecEnv(n, device=cuda_device, seed=3, mode="dp45")
    infl = torch.as_tensor(np.tile(g["influent"][:, None], (1, n))).to(cuda_device)
    obs_do, obs_ec = env.reset(influent=infl)
    assert obs_do.shape == (n, 9) and obs_ec.shape == (n, 9)
    assert np.allclose(obs_do[0].cpu().numpy(), g["reset_obs_do"], rtol=1e-5, atol=1e-7)
    total = torch.zeros(n, dtype=torch.float64, device=cuda_device)
    for k in range(463):
        a = torch.as_tensor(np.tile(g["action"][k][None, :], (n, 1))).to(cuda_device)
        (o_do, o_ec), state, reward, done, info = env.step(a)
        total += reward
        assert bool(done.all()) == bool(g["done"][k]) and bool(done.any()) == bool(g["done"][k])
    assert state.shape == (n, 15) and reward.shape == (n,)
    assert abs(float(total[0]) - g["reward"].sum()) < 1e-5 * abs(g["reward"].sum())
    assert torch.allclose(info["episode_return"], total, rtol=1e-12, atol=1e-14)
    assert float(info["episode_steps"][0]) == 463
    assert abs(float(info["Qw"][0]) - float(g["Qw"])) < 1e-5 * float(g["Qw"])
    # stepping finished envs: no-op
    (o_do, o_ec), state, reward, done, info = env.step(a)
    assert bool(done.all()) and float(reward.abs().max()) == 0.0
    assert int(info["status"].min()) == _abi.ST_DONE
    # autoreset: the next step restarts every finished env and steps it once
    env.autoreset = True
    (o_do, o_ec), state, reward, done, info = env.step(torch.as_tensor(
        np.tile(g["action"][0][None, :], (n, 1))).to(cuda_device))
    assert not bool(done.any()) and bool(info["restarted"].all())
    assert float(info["episode_steps"][0]) == 1
    assert abs(float(state[0, 0]) * 0.5 - g["t"][0]) < 1e-15


def test_autoreset_lockstep_path_draws_and_launches_nothing_between_episode_ends(built, cuda_device):
    """Envs started by one full reset end together: with autoreset the env must not draw influent or reset anything
    on the other steps (it did, every step, before), and must fall back to the masked path after a partial reset."""
    n = 64
    env = SbrOsVecEnv(n, device=cuda_device, seed=5, mode="dp45", autoreset=True)
    env.reset()
    a = torch.tensor([[2.0, 5.0]], dtype=torch.float64, device=cuda_device).repeat(n, 1)
    assert int(env.epoch.min()) == 1 == int(env.epoch.max())             # one influent draw so far
    for k in range(463):
        _, _, _, done, info = env.step(a)
        assert not bool(info["restarted"].any())
        assert bool(done.all()) == (k == 462)
        assert info["reset_obs"] is None
    assert int(env.epoch.max()) == 1                                     # nothing was drawn during the episode
    ret1 = info["episode_return"].clone()
    _, _, _, done, info = env.step(a)                                    # first step of the second episode
    assert bool(info["restarted"].all()) and not bool(done.any()) and float(info["episode_steps"].max()) == 1
    assert int(env.epoch.min()) == 2 == int(env.epoch.max())             # one influent draw per episode
    assert info["reset_obs"][0].shape == (n, 9) and bool(torch.isfinite(info["reset_obs"][0]).all())
    for k in range(462):
        _, _, _, done, info = env.step(a)
    assert bool(done.all()) and float(info["episode_steps"].min()) == 463
    assert bool(torch.isfinite(info["episode_return"]).all()) and not torch.equal(info["episode_return"], ret1)
    # a partial reset leaves lock-step: the general (masked, every step) path takes over and still restarts envs
    mask = torch.zeros(n, dtype=torch.bool, device=cuda_device)
    mask[: n // 2] = True
    env.reset(mask=mask)
    assert env._lockstep is False
    _, _, _, done, info = env.step(a)
    assert bool(info["restarted"][n // 2:].all()) and not bool(info["restarted"][: n // 2].any())
    assert float(info["episode_steps"].max()) == 1 and not bool(done.any())


def test_argument_errors_do_not_launch(built, cuda_device):
    lib = _abi.load()
    p, s = _abi.default_params(), schedule.os_schedule()
    import ctypes as C
    rc = lib.sbr_os_step(0, 0, None, None, C.byref(p), C.byref(s), None, None, None, None, None, None, None, 0, None,
                         None)
    assert rc == -1 and b"n must be positive" in lib.sbr_last_error()
    buf = core.OsBuffers(4, cuda_device)
    s.fill_pts = 0
    with pytest.raises(_abi.SbrLibraryError):
        core.os_reset(buf, torch.zeros((14, 4), dtype=torch.float64, device=cuda_device), p, s)


def test_full_size_properties_2p20(built, cuda_device):
    """BASELINE full size (2^20 envs on one GPU), 56 env.steps across the anoxic -> aerobic switch: determinism,
    batch-position invariance (the batch tiles 4096 distinct envs), agreement with the small batch, finite outputs."""
    from gym_sbr2_b200 import influent
    n, base = 1 << 20, 4096
    rng = np.random.RandomState(41)
    infl = np.stack([influent.mix_numpy(6, rng.randn(48)) for _ in range(64)], axis=1)
    infl = np.tile(infl, (1, base // 64)).copy()
    infl[0] = schedule.os_fill_flow(_abi.default_params().Qin)
    acts = [np.stack([8 * rng.rand(base), 15 * rng.rand(base)]) for _ in range(56)]
    rep = n // base
    p, s, tol = _abi.default_params(), schedule.os_schedule(), _abi.make_tol()
    dev = lambda a, r=1: torch.as_tensor(np.ascontiguousarray(np.tile(a, (1, r)))).to(cuda_device)

    def run(r):
        buf = core.OsBuffers(base * r, cuda_device)
        core.os_reset(buf, dev(infl, r), p, s, mode=_abi.MODE_DP45, tol=tol)
        for a in acts:
            core.os_step(buf, dev(a, r), p, s, mode=_abi.MODE_DP45, tol=tol)
        torch.cuda.synchronize()
        return buf
    big, small, again = run(rep), run(1), run(rep)
    st = big.st.view(_abi.OS_ROWS, rep, base)
    rows = [r for r in range(_abi.OS_ROWS) if r != _abi.OS_QW]                     # Qw is NaN until the episode ends
    assert bool((st[rows] == st[rows][:, :1]).all())
    assert torch.equal(big.st[rows][:, :base], small.st[rows]) and torch.equal(big.reward[:base], small.reward)
    assert torch.equal(big.st[rows], again.st[rows]) and torch.equal(big.state, again.state)
    assert bool(torch.isfinite(big.state).all()) and int(big.status.max()) == 0 and not bool(big.done.any())
    assert float(big.st[_abi.OS_STEPS].min()) == 56 == float(big.st[_abi.OS_STEPS].max())


def test_custom_start_state_against_oracle(built, cuda_device):
    """reset(x0=...) with a perturbed start state: fill solve + 30 steps against the oracle with the same x0."""
    from oracle import sbr_oracle as O
    g = load_episode("seed2_walk")
    rng = np.random.RandomState(6)
    x0 = np.array(O.X0_INIT) * np.exp(0.05 * rng.randn(14))
    x0[0] = O.X0_INIT[0]
    b = GpuOsBatch(1, cuda_device)
    od, oe = b.reset(g["influent"][:, None], x0=x0[:, None])
    o = O.SbrOsOracle()
    r_do, r_ec = o.reset(g["influent"], x0=x0)
    assert np.allclose(od[:, 0], r_do, rtol=1e-5, atol=1e-7) and np.allclose(oe[:, 0], r_ec, rtol=1e-5, atol=1e-7)
    for k in range(30):
        o_do, o_ec, st, r, done = b.step(g["action"][k][:, None])
        (_, _), r_st, r_r, _ = o.step(g["action"][k])
        ok, worst = parity.os_close(st[:, 0], r_st)
        assert ok, (k, worst)
        assert abs(r[0] - r_r) <= 1e-5 * abs(r_r) + parity.OS_REWARD_ATOL, k


def test_poisoned_actions_and_step_limit_are_flagged_not_fatal(built, cuda_device):
    """NaN / absurd actions and a starved step budget: the launch terminates, the affected env is flagged in
    `status`, its neighbours are untouched."""
    g = load_episode("seed0_const")
    n = 64
    infl = np.tile(g["influent"][:, None], (1, n))
    clean, dirty = GpuOsBatch(n, cuda_device), GpuOsBatch(n, cuda_device)
    clean.reset(infl); dirty.reset(infl)
    for k in range(60):
        a = np.tile(g["action"][k][:, None], (1, n))
        ref = clean.step(a)
        a[0, 5] = np.nan if k == 55 else a[0, 5]              # NaN DO set-point once, in the aerobic phase
        a[1, 9] = 1e300                                       # absurd NO3 set-point (clipped to 15 by the env)
        out = dirty.step(a)
    st = dirty.status
    assert st[5] & _abi.ST_NONFINITE
    keep = np.ones(n, dtype=bool); keep[[5, 9]] = False
    assert (st[keep] == 0).all()
    assert np.array_equal(out[2][:, keep], ref[2][:, keep]) and np.array_equal(out[3][keep], ref[3][keep])
    assert np.isfinite(out[2][:, 9]).all()
    starved = GpuOsBatch(4, cuda_device, tol=_abi.make_tol(1e-12, 1e-14, 1))
    starved.reset(infl[:, :4])
    starved.step(np.tile(g["action"][0][:, None], (1, 4)))
    assert (starved.status & _abi.ST_STEPLIMIT).all()


def _random_os_batch(n, device, seed):
    from gym_sbr2_b200 import influent
    rng = np.random.RandomState(seed)
    infl = np.stack([influent.mix_numpy(6, rng.randn(48)) for _ in range(64)], axis=1)
    infl = np.tile(infl, (1, (n + 63) // 64))[:, :n].copy()
    infl[0] = schedule.os_fill_flow(_abi.default_params().Qin)
    return torch.as_tensor(infl).to(device), rng


@pytest.mark.parametrize("mode", [_abi.MODE_RK4, _abi.MODE_DP45])
def test_k_step_launch_equals_k_single_steps(built, cuda_device, mode):
    """sbr_os_step_k: K consecutive env.steps in one launch (state in registers in between) == K launches of
    sbr_os_step, bit for bit: state rows (incl. the circular KLa history), rewards, final observation, counters summed.
    Covers blocks that cross the double-interval step 51, ragged batch sizes and the end of the episode."""
    n = 1000                                                       # 15 full tiles + a partial one
    infl, rng = _random_os_batch(n, cuda_device, 3)
    p, s, tol = _abi.default_params(), schedule.os_schedule(), _abi.make_tol()
    acts = torch.as_tensor(np.stack([np.stack([1 + 6 * rng.rand(n), 2 + 10 * rng.rand(n)]) for _ in range(463)])
                           ).to(cuda_device)
    one, blk = core.OsBuffers(n, cuda_device), core.OsBuffers(n, cuda_device)
    core.os_reset(one, infl, p, s, mode=mode, tol=tol)
    core.os_reset(blk, infl, p, s, mode=mode, tol=tol)
    k0 = 0
    for K in (1, 7, 16, 40, 200, 250):                             # 514 > 463: the last block runs past the episode end
        a = torch.zeros((K, 2, n), dtype=torch.float64, device=cuda_device)
        a[: min(K, 463 - k0)] = acts[k0:k0 + K]
        rew1 = torch.zeros((K, n), dtype=torch.float64, device=cuda_device)
        cnt = torch.zeros((2, n), dtype=torch.int64, device=cuda_device)
        stat = torch.zeros((n,), dtype=torch.int32, device=cuda_device)
        for k in range(K):
            was_done = one.done.clone().bool()
            core.os_step(one, a[k].contiguous(), p, s, mode=mode, tol=tol)
            rew1[k] = one.reward
            cnt += one.counters.to(torch.int64)
            stat |= torch.where(was_done & (k > 0), torch.zeros_like(stat), one.status)
            if k == 0:
                last = [one.obs_do.clone(), one.obs_ec.clone(), one.state.clone()]
            else:
                for dst, src in zip(last, (one.obs_do, one.obs_ec, one.state)):
                    dst.copy_(torch.where(was_done[None, :], dst, src))
        rewK = torch.full((K, n), 7.0, dtype=torch.float64, device=cuda_device)
        core.os_step(blk, a, p, s, mode=mode, tol=tol, rewards=rewK)
        rows = [r for r in range(_abi.OS_ROWS) if r != _abi.OS_QW]
        assert torch.equal(blk.st[rows].view(torch.int64), one.st[rows].view(torch.int64)), K
        assert torch.equal(rewK, rew1) and torch.equal(blk.done, one.done), K
        assert torch.equal(blk.obs_do, last[0]) and torch.equal(blk.obs_ec, last[1]) and torch.equal(blk.state, last[2]), K
        assert torch.equal(blk.counters.to(torch.int64), cnt) and torch.equal(blk.status, stat), K
        k0 += K
    assert bool(one.done.all()) and torch.equal(blk.st[_abi.OS_QW], one.st[_abi.OS_QW])
    assert float(one.st[_abi.OS_STEPS].min()) == 463 == float(blk.st[_abi.OS_STEPS].max())


def test_optional_outputs_and_unaligned_buffers(built, cuda_device):
    """obs_do / obs_ec / state may be NULL (not written, everything else identical); row strides and bases that rule
    out the 16-byte aligned bulk-copy staging take the plain-load path with identical results."""
    n = 200
    infl, rng = _random_os_batch(n, cuda_device, 8)
    p, s, tol = _abi.default_params(), schedule.os_schedule(), _abi.make_tol()
    acts = [torch.as_tensor(np.stack([1 + 6 * rng.rand(n), 2 + 10 * rng.rand(n)])).to(cuda_device) for _ in range(30)]
    full, lean = core.OsBuffers(n, cuda_device), core.OsBuffers(n, cuda_device)
    core.os_reset(full, infl, p, s, tol=tol); core.os_reset(lean, infl, p, s, tol=tol)
    lean.state.fill_(-5.0); lean.obs_ec.fill_(-6.0)
    for a in acts:
        core.os_step(full, a, p, s, tol=tol)
        core.os_step(lean, a, p, s, tol=tol, emit=("obs_do",))
    rows = [r for r in range(_abi.OS_ROWS) if r != _abi.OS_QW]
    assert torch.equal(lean.st[rows], full.st[rows]) and torch.equal(lean.obs_do, full.obs_do)
    assert torch.equal(lean.reward, full.reward)
    assert bool((lean.state == -5.0).all()) and bool((lean.obs_ec == -6.0).all())
    # odd row stride + bases offset by 8 bytes: strided views of wider buffers
    ld = n + 7
    odd = core.OsBuffers(ld + 1, cuda_device)
    view = core.OsBuffers.__new__(core.OsBuffers)
    for name in ("st", "obs_do", "obs_ec", "state", "counters"):
        setattr(view, name, getattr(odd, name)[:, 1:1 + n])
    view.st = odd.st[:, 1:1 + n]
    for name in ("reward", "done", "status"):
        setattr(view, name, torch.empty_like(getattr(full, name)))
    wide_infl = torch.zeros((14, ld + 1), dtype=torch.float64, device=cuda_device)
    wide_infl[:, 1:1 + n] = infl
    core.os_reset(view, wide_infl[:, 1:1 + n], p, s, tol=tol)
    wide_act = torch.zeros((2, ld + 1), dtype=torch.float64, device=cuda_device)
    for a in acts:
        wide_act[:, 1:1 + n] = a
        core.os_step(view, wide_act[:, 1:1 + n], p, s, tol=tol)
    assert torch.equal(view.st[rows], full.st[rows]) and torch.equal(view.state, full.state)
    assert torch.equal(view.reward, full.reward)


def test_trajectory_dump_matches_reference_trajectory(built, cuda_device):
    """SbrOS.trajectory() (gym_SBR_oneshot.py:1275-1288): the kernel's optional per-interval record against the
    unmodified reference's t_t / x_t / reward lists at the end of every env.step of a whole episode
    (oracle/make_golden_traj.py), and the reference's 18-tuple order through the single-env wrapper."""
    import os
    from conftest import GOLDEN
    from gym_sbr2_b200.envs.single import SbrOS
    g = np.load(os.path.join(GOLDEN, "sbros_v1_traj_seed0_const.npz"))
    env = SbrOS(device=cuda_device)
    np.random.seed(0)
    env.reset()
    assert np.array_equal(env.influent_mixed[1:], g["influent"][1:])       # same numpy stream as the reference
    done, steps = False, 0
    while not done:
        _, _, _, done, _ = env.step(g["action"])
        steps += 1
    assert steps == 463 == len(g["step_end_index"])
    tr = env._vec.trajectory(0)
    # 463 steps, three of them with two intervals, plus the post-draw and post-idle records, plus the post-fill state
    assert len(tr["t"]) == 1 + 466 + 2 and len(tr["step_end"]) == 463
    assert tr["t"][0] == g["t_fill_end"] and tr["t"][-1] == 0.5
    ok, worst = parity.os_close(tr["x"][0] / parity.STATE_SCALE, g["x_fill_end"] / parity.STATE_SCALE)
    assert ok, worst
    ends = tr["step_end"].copy()
    ends[-1] = len(tr["t"]) - 1                       # the reference's last step ends after settle + draw + idle
    assert np.allclose(tr["t"][ends], g["t_step_end"], rtol=0, atol=1e-12)
    ok, worst = parity.os_close(tr["x"][ends] / parity.STATE_SCALE, g["x_step_end"] / parity.STATE_SCALE)
    assert ok, worst
    for mine, ref in (("reward", "reward_t"), ("reward_EQI", "reward_EQI_t"), ("reward_OCI", "reward_OCI_t"),
                      ("reward_AE", "reward_AE_t"), ("reward_EC", "reward_EC_t")):
        assert np.allclose(tr[mine], g[ref], rtol=1e-5, atol=1e-9), mine
    assert np.array_equal(tr["u_do"][tr["step_end"] - 1], g["u_do_step_end"])
    assert np.array_equal(tr["u_ec"][tr["step_end"] - 1], g["u_ec_step_end"])
    tup = env.trajectory()
    assert len(tup) == 18 == len(g["tuple_names"]) and tup[9] is None and len(tup[12]) == 463
    assert np.array_equal(np.asarray(tup[5]), tr["x"][:, 8]) and len(tup[4]) == 463
    # the dump is off unless asked for, and it does not change the step
    plain = SbrOsVecEnv(1, device=cuda_device, mode="dp45")
    infl = torch.as_tensor(g["influent"])[:, None].to(cuda_device)
    plain.reset(influent=infl)
    assert plain.traj is None
    a = torch.as_tensor(g["action"])[None, :].to(cuda_device)
    for _ in range(463):
        plain.step(a)
    assert torch.equal(plain.buf.st[:14], env._vec.buf.st[:14])
    with pytest.raises(RuntimeError):
        plain.trajectory()
