"""Scratch: accuracy (vs a tight-tolerance run, in parity-tolerance units) and episode time of the interval-per-step
envs as a function of the DP45 tolerances."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_sbr2_b200.vec_env import SbrOsVecEnv, SbrV4VecEnv
from gym_sbr2_b200 import parity, _abi
dev = "cuda:0"
n = 1 << 18
scale = torch.as_tensor(parity.STATE_SCALE, device=dev, dtype=torch.float64)[:, None]
TOLS = [(1e-10, 1e-12), (1e-8, 1e-10), (1e-7, 1e-9), (3e-7, 3e-9), (1e-6, 1e-8)]


def units(x, ref, atol):
    return ((x - ref).abs() / (1e-5 * ref.abs() + atol * scale))


def run_os(tol, regime):
    env = SbrOsVecEnv(n, device=dev, seed=11, mode="dp45", rtol=tol[0], atol=tol[1])
    gen = torch.Generator(device=dev).manual_seed(5)
    if regime == "moderate":
        acts = [torch.stack([2 + torch.rand(n, dtype=torch.float64, device=dev, generator=gen),
                             4 + 2 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen)], dim=0).contiguous() for _ in range(8)]
    else:
        acts = [torch.stack([1 + 6 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen),
                             2 + 10 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen)], dim=0).contiguous() for _ in range(8)]
    env.reset()
    snaps = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(463):
        env.step_soa(acts[k % 8])
        if k in (50, 150, 275, 350, 461):
            snaps.append(env.buf.st[:14].clone())
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), snaps, env.buf.st[_abi.OS_RETURN].clone(), env.buf.st[10].clone()


def run_v4(tol):
    env = SbrV4VecEnv(n, device=dev, seed=3, mode="dp45", rtol=tol[0], atol=tol[1])
    gen = torch.Generator(device=dev).manual_seed(6)
    acts = [(0.6 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen) - 0.25) for _ in range(8)]
    env.reset()
    snaps = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(493):
        env.step_async(acts[k % 8])
        if k in (25, 100, 300, 491):
            snaps.append(env.buf.st[:14].clone())
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), snaps, env.buf.st[_abi.V4_RETURN].clone()


out = {}
for regime in ("moderate", "bench"):
    ref = None
    for tol in TOLS:
        ms, snaps, ret, snh = run_os(tol, regime)
        if ref is None:
            ref = (snaps, ret, snh); physical = snh > -0.5
            out["os_%s_truth_ms" % regime] = ms; out["os_%s_physical_frac" % regime] = float(physical.double().mean()); continue
        u = torch.stack([units(s, r, 1e-7).max(dim=0).values for s, r in zip(snaps, ref[0])]).max(dim=0).values
        u = u[physical]
        out["os_%s_rtol%g" % (regime, tol[0])] = dict(ms=ms, units_median=float(u.median()), units_p99=float(u.kthvalue(int(0.99 * u.numel())).values),
                                                     units_p999=float(u.kthvalue(int(0.999 * u.numel())).values), units_max=float(u.max()),
                                                     ret_absdiff_max=float((ret - ref[1])[physical].abs().max()))
ref = None
for tol in TOLS:
    ms, snaps, ret = run_v4(tol)
    if ref is None:
        ref = (snaps, ret); out["v4_truth_ms"] = ms; continue
    u = torch.stack([units(s, r, 1e-8).max(dim=0).values for s, r in zip(snaps, ref[0])]).max(dim=0).values
    out["v4_rtol%g" % tol[0]] = dict(ms=ms, units_median=float(u.median()), units_p99=float(u.kthvalue(int(0.99 * u.numel())).values),
                                    units_p999=float(u.kthvalue(int(0.999 * u.numel())).values), units_max=float(u.max()),
                                    ret_reldiff_max=float(((ret - ref[1]).abs() / ref[1].abs().clamp_min(1e-9)).max()))
print(json.dumps(out, indent=1))
