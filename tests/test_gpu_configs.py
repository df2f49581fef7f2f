"""BASELINE.json configs 3 and 4 (and the interval-path analogue of config 3) on the GPU, through the public vector
envs / C ABI: large batches with per-env influent and set-points, adaptive Dormand-Prince stepping, a random subset
replayed through the scipy oracle with the SAME inputs, plus size-independent properties over the whole batch."""
import numpy as np
import pytest
import torch

from gym_sbr2_b200 import _abi, core, parity, schedule
from gym_sbr2_b200.vec_env import SbrOsVecEnv, SbrV2VecEnv, X0_INIT
from oracle import sbr_oracle as O

pytestmark = pytest.mark.gpu


def test_config3_65536_envs_adaptive_steps_vs_oracle_and_rk4(built, cuda_device):
    """config 3: 65536 envs, per-env DO set-points and per-env stochastic influent, DP45 with per-env step control."""
    n = 65536
    env = SbrV2VecEnv(n, device=cuda_device, seed=33, mode="dp45", rtol=1e-8, atol=1e-10)
    env.reset()                                                    # per-env buffer_tank(0) draws on the device
    g = torch.Generator(device=cuda_device).manual_seed(4)
    action = torch.rand((n, 3), dtype=torch.float64, device=cuda_device, generator=g)
    obs, reward, done, info = env.step(action)
    x_dp = info["x_last"].clone()
    r_dp = reward.clone()
    cnt = info["counters"].to(torch.float64)
    assert int(info["status"].max()) == 0 and bool(done.all()) and bool(torch.isfinite(x_dp).all())
    # adaptive-step spread (the source of warp divergence): per-env RHS counts differ, rejects stay a small fraction
    rhs, rej = cnt[0], cnt[1]
    assert float(rhs.min()) > 528 * 7 and float(rhs.max()) < 19052 * 2
    assert float(rhs.std()) > 0 and float(rej.mean()) * 6 < 0.15 * float(rhs.mean())
    hist = torch.histc(rhs, bins=8, min=float(rhs.min()), max=float(rhs.max()))
    print("config3 RHS/env: min %.0f mean %.0f max %.0f rejects/env %.1f hist %s"
          % (rhs.min(), rhs.mean(), rhs.max(), rej.mean(), hist.to(torch.int64).tolist()))
    # the fixed-step mode on the same batch agrees far inside the parity tolerance for all but a handful of envs
    # that sit on a controller clamp / waste-layer switch (discontinuous dependence on the trajectory)
    env_rk = SbrV2VecEnv(n, device=cuda_device, mode="rk4")
    env_rk.reset(influent=env.influent)
    _, r_rk, _, info_rk = env_rk.step(action)
    bound = 1e-6 * info_rk["x_last"].abs() + 1e-10 * torch.as_tensor(parity.STATE_SCALE, device=cuda_device)[:, None]
    frac_out = float(((x_dp - info_rk["x_last"]).abs() > bound).any(dim=0).double().mean())
    assert frac_out < 2e-3, frac_out
    # a random subset through the scipy oracle with the same influent and action
    infl = env.influent.cpu().numpy()
    act = action.cpu().numpy()
    x_np, r_np, o_np = x_dp.cpu().numpy(), r_dp.cpu().numpy(), obs.cpu().numpy()
    # ... 32 random envs and the 32 slowest ones (largest RHS count: the envs the step controller worked hardest on)
    subset = np.concatenate([np.random.RandomState(1).choice(n, 32, replace=False),
                             torch.topk(rhs, 32).indices.cpu().numpy()])
    for i in subset:
        ref = O.sbr_v2_step(act[i], infl[:, i])
        ok, worst = parity.state_close(x_np[:, i], ref["x_last"])
        assert ok, (i, worst)
        if abs(ref["eff"][3] - 4) > 1e-3:
            assert abs(r_np[i] - ref["reward"]) <= 1e-5 * abs(ref["reward"]) + 1e-7, i
        assert np.allclose(o_np[i], ref["obs"], rtol=1e-5, atol=1e-8), i


def test_config4_stiff_stress_2p18_tight_tolerance_vs_tight_lsoda(built, cuda_device):
    """config 4: 2^18 envs started from states pushed toward high biomass / zero DO / high Ss (the fill and early
    anoxic phases are the stiffest, |lambda| ~ 7.5e3 1/d), DP45 at rtol 1e-9; a subset against LSODA at 1e-12."""
    n = 1 << 18
    rng = np.random.RandomState(8)
    base = 2048
    x0 = np.tile(np.array(X0_INIT)[:, None], (1, base))
    x0[5] *= rng.uniform(1.0, 1.3, base)            # Xbh
    x0[6] *= rng.uniform(1.0, 1.5, base)            # Xba
    x0[2] *= rng.uniform(1.0, 30.0, base)           # Ss
    x0[10] += rng.uniform(0.0, 5.0, base)           # Snh
    x0[8] = np.where(rng.rand(base) < 0.5, 0.0, x0[8])
    from gym_sbr2_b200 import influent
    infl = np.stack([influent.mix_numpy(0, rng.randn(48)) for _ in range(base)], axis=1)
    infl[0] = O.fill_flow()
    act = rng.rand(3, base)
    rep = n // base
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(np.tile(a, (1, rep)))).to(cuda_device)
    p, s = _abi.default_params(), schedule.cycle_schedule()
    tol = _abi.make_tol(1e-9, 1e-11, 20000)
    out = core.cycle_v2(dev(x0), dev(infl), dev(act), p, s, mode=_abi.MODE_DP45, tol=tol)
    torch.cuda.synchronize()
    st = out.status.cpu().numpy()
    assert not (st & (_abi.ST_NONFINITE | _abi.ST_STEPLIMIT)).any()
    xl = out.x_last.cpu().numpy()
    assert np.array_equal(xl[:, :base], xl[:, -base:])                       # batch-position invariance
    cnt = out.counters.to(torch.float64)
    print("config4 RHS/env: mean %.0f max %.0f rejects/env %.1f" % (cnt[0].mean(), cnt[0].max(), cnt[1].mean()))
    kw = dict(rtol=1e-12, atol=1e-12, mxstep=50000)
    checked = 0
    slowest = torch.topk(cnt[0][:base], 32).indices.cpu().numpy()           # the stiffest starts of the batch
    for i in np.concatenate([rng.choice(base, 32, replace=False), slowest]):
        ref = O.sbr_v2_step(act[:, i], infl[:, i], x0=x0[:, i], ode_kw=kw)
        assert (st[i] != 0) == (ref["status"] != 0), i
        if st[i] == 0:
            ok, worst = parity.state_close(xl[:, i], ref["x_last"], rtol=1e-6, atol_frac=1e-10)
            assert ok, (i, worst)
            checked += 1
    assert checked >= 48, checked


def test_interval_path_65536_envs_subset_vs_oracle(built, cuda_device):
    """Interval-per-step analogue of config 3: 65536 SBROS-v1 envs, per-env influent (buffer_tank(6) draws) and
    slowly varying per-env set-points, 140 env.steps across the anoxic -> aerobic switch; 8 envs replayed through the
    oracle with the same influent and actions."""
    n, steps = 65536, 140
    env = SbrOsVecEnv(n, device=cuda_device, seed=12, mode="dp45")
    obs_do, obs_ec = env.reset()
    pick = np.random.RandomState(2).choice(n, 8, replace=False)
    infl = env.influent.cpu().numpy()[:, pick]
    oracles = []
    for j in range(len(pick)):
        o = O.SbrOsOracle()
        od, oe = o.reset(infl[:, j])
        assert np.allclose(obs_do[pick[j]].cpu().numpy(), od, rtol=1e-5, atol=1e-7)
        assert np.allclose(obs_ec[pick[j]].cpu().numpy(), oe, rtol=1e-5, atol=1e-7)
        oracles.append(o)
    g = torch.Generator(device=cuda_device).manual_seed(9)
    a = torch.stack([1 + 5 * torch.rand(n, dtype=torch.float64, device=cuda_device, generator=g),
                     3 + 8 * torch.rand(n, dtype=torch.float64, device=cuda_device, generator=g)], dim=1)
    for k in range(steps):
        a = a + 0.05 * torch.randn(a.shape, dtype=torch.float64, device=cuda_device, generator=g)
        a[:, 0].clamp_(0.5, 7.0); a[:, 1].clamp_(1.0, 14.0)
        (o_do, o_ec), state, reward, done, info = env.step(a)
        assert not bool(done.any())
        a_np = a[pick].cpu().numpy()
        st_np, r_np = state[pick].cpu().numpy(), reward[pick].cpu().numpy()
        od_np, oe_np = o_do[pick].cpu().numpy(), o_ec[pick].cpu().numpy()
        for j, o in enumerate(oracles):
            (r_do, r_ec), r_st, r_r, r_done = o.step(a_np[j])
            ok, worst = parity.os_close(st_np[j], r_st)
            assert ok, (k, j, worst)
            assert parity.os_obs_close(od_np[j], r_do, r_st, "do")[0], (k, j)
            assert parity.os_obs_close(oe_np[j], r_ec, r_st, "ec")[0], (k, j)
            assert abs(r_np[j] - r_r) <= 1e-5 * abs(r_r) + parity.OS_REWARD_ATOL, (k, j)
    assert int(info["status"].max()) == 0
    assert float(info["episode_steps"].min()) == steps
    cnt = info["counters"].to(torch.float64)
    print("interval path RHS/env-step: mean %.1f max %.0f" % (cnt[0].mean(), cnt[0].max()))


def test_wide_setpoints_step_limit_flags_are_the_unphysical_envs(built, cuda_device):
    """bench.py's stress case (DO set-point U(1,7), NO3 set-point U(2,12) per step): heavy carbon dosing drives some envs
    into negative ammonia, towards the pole of Snh / (Knh + Snh) -- the regime where the reference's own LSODA gives up
    ("excess work").  There the adaptive stepper runs into max_steps inside one interval and says so: the flagged envs
    carry SBR_ST_STEPLIMIT and nothing else, every one of them has left the physical regime (Snh < 0), and with a
    larger step budget the same episode finishes without a flag."""
    n = 16384
    def run(max_steps):
        env = SbrOsVecEnv(n, device=cuda_device, seed=77, mode="dp45", max_steps=max_steps)
        env.reset()
        gen = torch.Generator(device=cuda_device).manual_seed(5)
        acts = [torch.stack([1 + 6 * torch.rand(n, dtype=torch.float64, device=cuda_device, generator=gen),
                             2 + 10 * torch.rand(n, dtype=torch.float64, device=cuda_device, generator=gen)], dim=1)
                for _ in range(8)]
        flagged = torch.zeros(n, dtype=torch.int32, device=cuda_device)
        min_snh = torch.full((n,), 1e9, dtype=torch.float64, device=cuda_device)
        for k in range(463):
            env.step_async(acts[k % 8])
            flagged |= env.buf.status
            min_snh = torch.minimum(min_snh, env.buf.st[10])
        return flagged, min_snh, env
    flagged, min_snh, env = run(200)
    bad = flagged != 0
    print("stress: %d of %d envs flagged, bits %s, min Snh of flagged envs in [%.3f, %.3f], %d unflagged envs below 0"
          % (int(bad.sum()), n, sorted(set(flagged[bad].cpu().tolist())), float(min_snh[bad].min()) if bool(bad.any()) else 0,
             float(min_snh[bad].max()) if bool(bad.any()) else 0, int(((min_snh < 0) & ~bad).sum())))
    assert bool(bad.any()) and bool((flagged[bad] == _abi.ST_STEPLIMIT).all())
    assert bool((min_snh[bad] < 0).all())                                    # flagged => unphysical
    assert bool(env.buf.done.all()) and bool(torch.isfinite(env.buf.st[:14]).all())
    flagged2, _, _ = run(20000)
    assert int((flagged2 & _abi.ST_STEPLIMIT).sum()) == 0
