"""GPU parity tests of the cycle-per-step path (SBR-v2) -- all calls go through the C ABI (ctypes)."""
import ctypes as C

import numpy as np
import pytest
import torch

from gym_sbr2_b200 import _abi, core, parity, schedule
from gym_sbr2_b200.vec_env import SbrV2VecEnv, X0_INIT
from oracle import sbr_oracle as O
from oracle.twin import binding as twin

pytestmark = pytest.mark.gpu


def _dev(a, device):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(device)


def _golden_inputs(g, device):
    n = len(g["seed"])
    x0 = np.tile(np.array(O.X0_INIT, dtype=float)[:, None], (1, n))
    infl = g["influent"].T.copy()
    infl[0] = O.fill_flow()
    return _dev(x0, device), _dev(infl, device), _dev(g["action"].T, device)


def test_library_loaded_and_device_visible(built, cuda_device):
    assert _abi.load().sbr_device_count() >= 1


def test_rhs_kernel_matches_reference_samples(built, cuda_device, stage_samples):
    s = stage_samples
    p = _abi.default_params()
    x = _dev(s["x"].T, cuda_device)
    n = x.shape[1]
    load = _dev(np.tile(s["load"][:, None], (1, n)), cuda_device)
    kla, ec = _dev(s["kla"], cuda_device), _dev(s["ec"], cuda_device)
    for tail, ref in ((_abi.TAIL_REACT, s["d_react"]), (_abi.TAIL_FILL, s["d_fill"]), (_abi.TAIL_EC, s["d_ec"])):
        dx = core.rhs(x, kla, p, tail, ec=ec, loading=load).cpu().numpy().T
        scale = np.abs(ref).max(axis=1, keepdims=True)
        # the three reciprocals of the RHS are MUFU.RCP64H + one Newton step (2 DFMA): relative error <= ~2^-36;
        # measured RHS error 2e-11 (cubic correction, -DSBR_RCP_NEWTON=3: 2.5e-14), see sbr_core.cuh rcp()
        assert np.all(np.abs(dx - ref) <= 1e-10 * np.abs(ref) + 1e-10 * scale), np.abs(dx - ref).max()


@pytest.mark.parametrize("mode", [_abi.MODE_RK4, _abi.MODE_DP45])
def test_cycle_kernel_matches_reference_golden(built, cuda_device, golden_v2, mode):
    g = golden_v2
    x0, infl, act = _golden_inputs(g, cuda_device)
    out = core.cycle_v2(x0, infl, act, _abi.default_params(), schedule.cycle_schedule(), mode=mode)
    torch.cuda.synchronize()
    assert int(out.status.max()) == 0
    ok, worst = parity.state_close(out.x_last.cpu().numpy().T, g["x_last"])
    assert ok, worst
    assert np.allclose(out.reward.cpu().numpy(), g["reward"], rtol=1e-5, atol=1e-7)
    assert np.allclose(out.obs.cpu().numpy().T, g["obs"], rtol=1e-5, atol=1e-8)
    aux = out.aux.cpu().numpy()
    assert np.allclose(aux[_abi.AUX_NAMES.index("Qw")], g["Qw"], rtol=1e-5, atol=1e-9)
    assert np.allclose(aux[_abi.AUX_NAMES.index("EQI")], g["EQI"], rtol=1e-5)
    assert np.allclose(aux[3:9].T, g["eff"], rtol=1e-5, atol=1e-8)
    for name in ("kla3_mean", "kla5_mean", "kla8_mean"):
        ok, worst = parity.scalar_close(aux[_abi.AUX_NAMES.index(name)], g[name], 4 * parity.KLA_SCALE)
        assert ok, (name, worst)


@pytest.mark.parametrize("mode", [_abi.MODE_RK4, _abi.MODE_DP45])
def test_cycle_kernel_closer_to_converged_than_reference(built, cuda_device, golden_v2_tight, mode):
    gt = golden_v2_tight
    x0, infl, act = _golden_inputs(gt, cuda_device)
    out = core.cycle_v2(x0, infl, act, _abi.default_params(), schedule.cycle_schedule(), mode=mode)
    xl = out.x_last.cpu().numpy()
    for j in range(len(gt["seed"])):
        assert np.abs(xl[:, j] / gt["x_last"][j] - 1).max() < 5e-7
    assert np.abs(out.reward.cpu().numpy() - gt["reward"]).max() < 1e-7


def _random_batch(n, seed, device, spread=0.05):
    rng = np.random.RandomState(seed)
    x0 = np.array(X0_INIT)[:, None] * np.exp(spread * rng.randn(14, n))
    x0[0] = X0_INIT[0]
    from gym_sbr2_b200 import influent
    infl = np.stack([influent.mix_numpy(0, rng.randn(48)) for _ in range(min(n, 256))], axis=1)
    infl = np.tile(infl, (1, (n + infl.shape[1] - 1) // infl.shape[1]))[:, :n].copy()
    infl[0] = O.fill_flow()
    act = rng.rand(3, n)
    return x0, infl, act


@pytest.mark.parametrize("mode,tol", [(_abi.MODE_RK4, 2e-10), (_abi.MODE_DP45, 2e-6)])
def test_4096_envs_every_env_against_cpu_twin(built, cuda_device, mode, tol):
    """BASELINE config 2 (4096 envs, random set-points): every env against the g++ build of the same stepper.
    RK4 differs only by FMA contraction / reciprocal rounding; DP45 may take different accept/reject decisions at
    the margin, so it agrees at tolerance level."""
    n = 4096
    x0, infl, act = _random_batch(n, 11, cuda_device)
    p, s = _abi.default_params(), schedule.cycle_schedule()
    out = core.cycle_v2(_dev(x0, cuda_device), _dev(infl, cuda_device), _dev(act, cuda_device), p, s, mode=mode)
    ref = twin.cycle_v2(x0, infl, act, p, s, mode=mode)
    xl = out.x_last.cpu().numpy()
    assert np.array_equal(out.status.cpu().numpy(), ref["status"])
    ok, worst = parity.state_close(xl.T, ref["x_last"].T, rtol=tol, atol_frac=tol * 1e-3)
    assert ok, worst
    same = np.abs(ref["aux"][6] - 4) > 1e-6          # away from the Snh = 4 reward step
    assert np.allclose(out.reward.cpu().numpy()[same], ref["reward"][same], rtol=max(tol, 1e-9) * 10, atol=1e-9)
    cnt = out.counters.cpu().numpy().astype(np.float64)
    if mode == _abi.MODE_RK4:
        assert np.array_equal(cnt.astype(np.uint32), ref["counters"])
    else:
        # tripwire for the adaptive stepper's bookkeeping (a device build once broke the first-same-as-last carry:
        # 50x the rejected steps, 3x the RHS evaluations, see the note in pid_phase): same work as the twin within
        # the noise of marginal accept/reject decisions
        rhs_ref, rej_ref = ref["counters"][0].astype(np.float64), ref["counters"][1].astype(np.float64)
        assert abs(cnt[0].mean() - rhs_ref.mean()) < 0.02 * rhs_ref.mean(), (cnt[0].mean(), rhs_ref.mean())
        assert cnt[1].mean() < 1.5 * rej_ref.mean() + 5, (cnt[1].mean(), rej_ref.mean())


def test_4096_envs_subset_against_scipy_oracle(built, cuda_device):
    """BASELINE config 2: a random subset replayed through the oracle (scipy LSODA) with the same inputs."""
    n = 4096
    x0, infl, act = _random_batch(n, 12, cuda_device, spread=0.0)
    p, s = _abi.default_params(), schedule.cycle_schedule()
    out = core.cycle_v2(_dev(x0, cuda_device), _dev(infl, cuda_device), _dev(act, cuda_device), p, s)
    xl, rw, ob = out.x_last.cpu().numpy(), out.reward.cpu().numpy(), out.obs.cpu().numpy()
    aux = out.aux.cpu().numpy()
    rng = np.random.RandomState(0)
    for i in rng.choice(n, 64, replace=False):
        ref = O.sbr_v2_step(act[:, i], infl[:, i])
        ok, worst = parity.state_close(xl[:, i], ref["x_last"])
        assert ok, (i, worst)
        if abs(ref["eff"][3] - 4) > 1e-3:
            assert abs(rw[i] - ref["reward"]) <= 1e-5 * abs(ref["reward"]) + 1e-7, i
        assert np.allclose(ob[:, i], ref["obs"], rtol=1e-5, atol=1e-8), i
        assert np.isclose(aux[1, i], ref["Qw"], rtol=1e-5, atol=1e-9), i


def test_vec_env_reset_and_step_semantics(built, cuda_device, golden_v2):
    """Gym surface: reset obs = reference's (sum of x0 and influent), step tuple shapes, done always True."""
    g = golden_v2
    n = len(g["seed"])
    env = SbrV2VecEnv(n, device=cuda_device, seed=0)
    obs0 = env.reset(influent=_dev(g["influent"].T, cuda_device))
    assert obs0.shape == (n, 3)
    assert np.allclose(obs0.cpu().numpy(), g["reset_obs"], rtol=1e-14, atol=0)
    obs, reward, done, info = env.step(_dev(g["action"], cuda_device))
    assert obs.shape == (n, 3) and reward.shape == (n,) and done.dtype == torch.bool and bool(done.all())
    assert np.allclose(reward.cpu().numpy(), g["reward"], rtol=1e-5, atol=1e-7)
    assert np.allclose(obs.cpu().numpy(), g["obs"], rtol=1e-5, atol=1e-8)
    # like the reference, a second step replays the cycle from x0_init (gym_SBR_env2.py:88-99)
    obs2, reward2, _, _ = env.step(_dev(g["action"], cuda_device))
    assert torch.equal(reward2, reward) and torch.equal(obs2, obs)
    # default reset draws a fresh influent per env on the device
    o = env.reset()
    assert o.shape == (n, 3) and bool(torch.isfinite(o).all()) and float(o[:, 2].std()) > 0


def test_ragged_sizes_strided_views_and_batch_invariance(built, cuda_device):
    """n = 1, n not a multiple of the block, ld > n; an env's result must not depend on batch size/position."""
    p, s = _abi.default_params(), schedule.cycle_schedule()
    x0, infl, act = _random_batch(333, 3, cuda_device)
    full = core.cycle_v2(_dev(x0, cuda_device), _dev(infl, cuda_device), _dev(act, cuda_device), p, s)
    xl_full = full.x_last.clone()
    rw_full = full.reward.clone()
    for lo, hi in ((0, 1), (5, 70), (100, 333)):
        n = hi - lo
        bx, bi, ba = _dev(x0, cuda_device)[:, lo:hi], _dev(infl, cuda_device)[:, lo:hi], _dev(act, cuda_device)[:, lo:hi]
        out = core.CycleV2Out(333, cuda_device)
        for name in ("x_last", "obs", "aux", "counters"):
            setattr(out, name, getattr(out, name)[:, :n])
        out.reward, out.status = out.reward[:n], out.status[:n]
        core.cycle_v2(bx, bi, ba, p, s, out=out)          # ld = 333 > n: strided SoA views
        assert torch.equal(out.x_last, xl_full[:, lo:hi])
        assert torch.equal(out.reward, rw_full[lo:hi])


@pytest.mark.parametrize("mode", [_abi.MODE_RK4, _abi.MODE_DP45])
def test_env_ordering_does_not_change_results(built, cuda_device, mode):
    """perm: thread i works on env perm[i] (divergence-aware ordering).  Every env's result is bit-identical to the
    identity mapping, buffers keep the caller's env order; SbrV2VecEnv(order='action') sorts by the first set-point."""
    n = 3000
    x0, infl, act = _random_batch(n, 9, cuda_device)
    p, s = _abi.default_params(), schedule.cycle_schedule()
    X0, IN, AC = _dev(x0, cuda_device), _dev(infl, cuda_device), _dev(act, cuda_device)
    base = core.cycle_v2(X0, IN, AC, p, s, mode=mode)
    xl, rw, cn = base.x_last.clone(), base.reward.clone(), base.counters.clone()
    for perm in (torch.argsort(AC[0]), torch.randperm(n, device=cuda_device), torch.arange(n - 1, -1, -1, device=cuda_device)):
        out = core.cycle_v2(X0, IN, AC, p, s, mode=mode, perm=perm.to(torch.int64).contiguous())
        assert torch.equal(out.x_last, xl) and torch.equal(out.reward, rw) and torch.equal(out.counters, cn)
    env_a = SbrV2VecEnv(n, device=cuda_device, mode="dp45", order="action")
    env_n = SbrV2VecEnv(n, device=cuda_device, mode="dp45", order="none")
    assert env_a.order == "action" and SbrV2VecEnv(8, device=cuda_device, mode="rk4").order == "none"
    infl_t = torch.as_tensor(infl).to(cuda_device)
    env_a.reset(influent=infl_t); env_n.reset(influent=infl_t)
    a = AC.t().contiguous()
    ra, rn = env_a.step(a)[1].clone(), env_n.step(a)[1].clone()
    assert torch.equal(ra, rn)


def test_status_flags_and_error_returns(built, cuda_device):
    p, s = _abi.default_params(), schedule.cycle_schedule()
    x0, infl, act = _random_batch(64, 4, cuda_device)
    x0[8, 3] = np.nan                       # poisoned env -> flagged, neighbours untouched
    out = core.cycle_v2(_dev(x0, cuda_device), _dev(infl, cuda_device), _dev(act, cuda_device), p, s)
    st = out.status.cpu().numpy()
    assert st[3] & _abi.ST_NONFINITE and (np.delete(st, 3) == 0).all()
    assert bool(torch.isfinite(out.reward[:3]).all())
    with pytest.raises(_abi.SbrLibraryError):
        core.cycle_v2(_dev(x0, cuda_device).cpu(), _dev(infl, cuda_device), _dev(act, cuda_device), p, s)
    with pytest.raises(_abi.SbrLibraryError):
        core.cycle_v2(_dev(x0, cuda_device), _dev(infl, cuda_device), _dev(act, cuda_device), p, s, mode=7)
    bad = schedule.cycle_schedule()
    bad.n_int[2] = 0
    with pytest.raises(_abi.SbrLibraryError):
        core.cycle_v2(_dev(x0, cuda_device), _dev(infl, cuda_device), _dev(act, cuda_device), p, bad)


def test_single_interval_kernel_against_twin(built, cuda_device, stage_samples):
    s = stage_samples
    p = _abi.default_params()
    rng = np.random.RandomState(5)
    n = 200
    x = np.array(O.X0_INIT)[:, None] * np.exp(0.1 * rng.randn(14, n))
    kla, ec = 240 * rng.rand(n), 0.0005 * rng.rand(n)
    load = np.tile(s["load"][:, None], (1, n))
    for tail in (_abi.TAIL_REACT, _abi.TAIL_FILL, _abi.TAIL_EC):
        for mode, tol in ((_abi.MODE_RK4, 1e-12), (_abi.MODE_DP45, 1e-7)):
            ref, _ = twin.integrate_interval(x, kla, p, tail, 0.02 / 24, 10, mode=mode, ec=ec, loading=load)
            xd = _dev(x, cuda_device)
            core.integrate_interval(xd, _dev(kla, cuda_device), p, tail, 0.02 / 24, 10, mode=mode,
                                    ec=_dev(ec, cuda_device), loading=_dev(load, cuda_device))
            ok, worst = parity.state_close(xd.cpu().numpy().T, ref.T, rtol=tol, atol_frac=tol * 1e-3)
            assert ok, (tail, mode, worst)


def test_full_size_properties_2p20(built, cuda_device):
    """BASELINE full size (2^20 envs on one GPU): size-independent properties -- determinism, batch-position
    invariance (the batch tiles 4096 distinct envs), finite outputs, exact discrete outputs."""
    n = 1 << 20
    base = 4096
    x0, infl, act = _random_batch(base, 21, cuda_device)
    rep = n // base
    X0, IN, AC = (_dev(np.tile(a, (1, rep)), cuda_device) for a in (x0, infl, act))
    p, s = _abi.default_params(), schedule.cycle_schedule()
    out = core.cycle_v2(X0, IN, AC, p, s)
    xl = out.x_last.clone()
    rw = out.reward.clone()
    assert int(out.status.max()) == 0 and bool(torch.isfinite(xl).all())
    assert torch.equal(xl.view(14, rep, base)[:, 0], xl.view(14, rep, base)[:, rep - 1])
    assert bool((xl.view(14, rep, base) == xl.view(14, rep, base)[:, :1]).all())
    small = core.cycle_v2(_dev(x0, cuda_device), _dev(infl, cuda_device), _dev(act, cuda_device), p, s)
    assert torch.equal(small.x_last, xl[:, :base]) and torch.equal(small.reward, rw[:base])
    again = core.cycle_v2(X0, IN, AC, p, s)
    assert torch.equal(again.x_last, xl) and torch.equal(again.reward, rw)
    assert bool((out.counters[0] == 19052).all())


def test_fp64_probe_runs(built, cuda_device):
    sink, flops = core.fp64_probe(148 * 8, 256, 2000, cuda_device)
    torch.cuda.synchronize()
    assert flops == 2.0 * 8 * 2000 * 148 * 8 * 256 and bool(torch.isfinite(sink).all())


def test_settle_and_draw_stage_matches_reference_samples(built, cuda_device, stage_samples):
    """A7 / A8 (and B5) at stage level: the settler's closed form and the draw / waste / effluent-quality arithmetic
    through sbr_settle_draw against samples recorded from the unmodified reference's sim_settling (two 10-layer odeint
    solves) and sim_drawing + cal_eq (sub_phases_FB.py:716-915), tests/golden/stage_samples.npz."""
    s = stage_samples
    n = len(s["settle_x"])
    x = core.soa1(torch.as_tensor(np.ascontiguousarray(s["settle_x"].T)).to(cuda_device))
    sX, out, status = core.settle_draw(x, _abi.default_params(), schedule.cycle_schedule().settle_time)
    assert int(status.abs().max()) == 0
    sX, out, x7 = sX.cpu().numpy().T, out.cpu().numpy().T, x.cpu().numpy().T
    # the layer solids come from the reference's LSODA solve (rtol 1.49e-8): the closed form is the exact solution
    assert np.allclose(sX, s["settle_sX"], rtol=2e-7, atol=1e-6 * s["settle_Xf"][:, None])
    assert np.array_equal(out[:, 0], s["settle_Xf"])
    for i in range(n):
        ok, worst = parity.state_close(x7[i], s["draw_x7"][i])
        assert ok, (i, worst)
    assert np.allclose(out[:, 1], s["draw_Qw"], rtol=1e-5) and np.allclose(out[:, 2], s["draw_EQI"], rtol=1e-5)
    assert np.allclose(out[:, 3:], s["draw_eff"], rtol=1e-5, atol=1e-9)
    with pytest.raises(_abi.SbrLibraryError):
        core.settle_draw(x, _abi.default_params(), 0.0)


@pytest.mark.parametrize("mode", ["rk4", "dp45"])
def test_config1_4096_envs_random_kla_actions(built, cuda_device, mode):
    """BASELINE configs[1] as written: 4096 vectorised envs, random KLa actions (PID bypassed, action_kind='kla'), one
    cycle each: every env against the CPU twin, 64 of them (the largest KLa sums among them) against the scipy oracle
    (LSODA at 1e-12: open loop, the default-tolerance LSODA is itself 1.5 tolerance units from its converged run)."""
    TIGHT = dict(rtol=1e-12, atol=1e-12, mxstep=50000)
    n = 4096
    env = SbrV2VecEnv(n, device=cuda_device, seed=21, mode=mode, action_kind="kla")
    env.reset()
    g = torch.Generator(device=cuda_device).manual_seed(3)
    action = torch.rand((n, 3), dtype=torch.float64, device=cuda_device, generator=g)
    obs, reward, done, info = env.step(action)
    assert int(info["status"].max()) == 0 and bool(done.all())
    kla3 = info["kla3_mean"].cpu().numpy()
    act = action.cpu().numpy()
    assert np.allclose(kla3, 240 * act[:, 0], rtol=1e-12)                      # constant over phase 3: mean == value
    x, r = info["x_last"].cpu().numpy(), reward.cpu().numpy()
    ld = env._loading.cpu().numpy()
    t = twin.cycle_v2(env.x0.cpu().numpy(), ld, act.T.copy(), twin.default_params(), env.sched,
                      mode=env.mode, tol=env.tol)
    bound = 1e-5 * np.abs(t["x_last"]) + 1e-9 * parity.STATE_SCALE[:, None]
    assert (np.abs(x - t["x_last"]) <= bound * (0.05 if mode == "dp45" else 1e-4)).all()
    infl = env.influent.cpu().numpy()
    pick = np.concatenate([np.random.RandomState(0).choice(n, 48, replace=False), np.argsort(-act.sum(axis=1))[:16]])
    for i in pick:
        ref = O.sbr_v2_step(act[i], infl[:, i], raw_kla=True, ode_kw=TIGHT)
        ok, worst = parity.state_close(x[:, i], ref["x_last"])
        assert ok, (i, worst)
        if abs(ref["eff"][3] - 4) > 1e-3:
            assert abs(r[i] - ref["reward"]) <= 1e-5 * abs(ref["reward"]) + 1e-7, i
    with pytest.raises(ValueError):
        SbrV2VecEnv(8, device=cuda_device, action_kind="raw")


def test_cycle_trajectory_kernel_matches_reference_run_outputs(built, cuda_device):
    """sbr_cycle_v2_traj through the C ABI and SbrV2VecEnv.trajectory(): state at the end of every PID interval, post-draw
    state and per-interval KLa against SBR_model_FB.run's `t`, `x`, kla3 / kla5 / kla8 of the UNMODIFIED reference
    (tests/golden/sbr_v2_traj_seed0.npz); tolerances as argued in tests/test_twin_parity.py."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sbr_v2_traj_seed0.npz"))
    n = 5
    infl = torch.as_tensor(np.tile(g["influent"][:, None], (1, n)), dtype=torch.float64, device=cuda_device)
    act = torch.as_tensor(np.tile(g["action"][None, :], (n, 1)), dtype=torch.float64, device=cuda_device)
    for mode, kw, ref, rtol, floor in (("dp45", dict(rtol=1e-9, atol=1e-11), "tight", 1e-6, 1e-9),
                                       ("dp45", dict(rtol=1e-9, atol=1e-11), "default", 1e-5, 1e-7),
                                       ("rk4", {}, "default", 1e-5, 1e-6)):
        env = SbrV2VecEnv(n, device=cuda_device, seed=0, mode=mode, **kw)
        env.reset(influent=infl)
        tr = env.trajectory(act)
        t, x, kla = tr["t"].cpu().numpy(), tr["x"].cpu().numpy(), tr["kla"].cpu().numpy()
        assert t.shape == (529, n) and x.shape == (529, 14, n) and not np.isnan(x).any()
        assert np.array_equal(x[:, :, 0], x[:, :, n - 1])
        ends_x = np.concatenate([x[:492, :, 2], x[493:, :, 2]])
        ends_t, ends_k = np.concatenate([t[:492, 2], t[493:, 2]]), np.concatenate([kla[:492, 2], kla[493:, 2]])
        assert np.allclose(ends_t, g[ref + "_t"], rtol=1e-12, atol=1e-15)
        want = g[ref + "_x"].copy()
        if mode == "rk4":
            assert np.allclose(ends_x[:4, 8], want[:4, 8], rtol=2e-3, atol=5e-6)
            ends_x[:4, 8] = want[:4, 8]
        ok, worst = parity.state_close(ends_x, want, rtol=rtol, atol_frac=floor)
        assert ok, (mode, ref, worst)
        assert parity.state_close(x[492, :, 2], g[ref + "_x_post_draw"], rtol=rtol, atol_frac=floor)[0]
        for name, lo, hi in (("kla3", 72, 295), ("kla5", 481, 492), ("kla8", 492, 528)):
            assert np.allclose(ends_k[lo:hi], g[ref + "_" + name], rtol=1e-5, atol=2e-4), (mode, name)
        # same end state and reward as the timed step of the same env
        obs, reward, done, info = env.step(act)
        assert parity.state_close(tr["x_last"].cpu().numpy().T, info["x_last"].cpu().numpy().T, rtol=1e-7)[0]
        assert np.allclose(tr["reward"].cpu().numpy(), reward.cpu().numpy(), rtol=1e-7)
