"""Scratch measurement (not the graded bench): kernel time of both paths in every integrator mode."""
import json
import sys
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from gym_sbr2_b200 import _abi, core, schedule
from gym_sbr2_b200.vec_env import SbrV2VecEnv, SbrOsVecEnv


def timed(fn, reps=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    dev = torch.device("cuda:0")
    out = {}
    gen = torch.Generator(device=dev).manual_seed(5)
    action = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=gen)
    for name, mode, kw, nsub in (("A_rk4_ref", "rk4", {}, None), ("A_rk4_sub5", "rk4", {}, 5), ("A_rk4_sub4", "rk4", {}, 4),
                                 ("A_dp45_1e-6", "dp45", dict(rtol=1e-6, atol=1e-8), None),
                                 ("A_dp45_1e-7", "dp45", dict(rtol=1e-7, atol=1e-9), None),
                                 ("A_dp45_1e-8", "dp45", dict(rtol=1e-8, atol=1e-10), None)):
        env = SbrV2VecEnv(n, device=dev, seed=1, mode=mode, **kw)
        if nsub:
            for k in range(8):
                if env.sched.n_sub[k]:
                    env.sched.n_sub[k] = nsub
        env.reset()
        ms = timed(lambda: env.step_async(action))
        o = env._out
        cnt = o.counters.to(torch.float64)
        out[name] = dict(ms=ms, cycle_steps_per_s=n / ms * 1e3, rhs_mean=float(cnt[0].mean()), rhs_max=float(cnt[0].max()),
                         rej_mean=float(cnt[1].mean()), bad=int((o.status != 0).sum()))
        if name == "A_rk4_ref":
            ref_x = o.x_last.clone(); ref_r = o.reward.clone()
        else:
            rel = ((o.x_last - ref_x).abs() / (1e-5 * ref_x.abs() + 1e-9 * torch.tensor(
                [1.32, 30, 30, 1500, 150, 3000, 2000, 600, 8, 20, 20, 10, 10, 10], device=dev, dtype=torch.float64)[:, None]))
            out[name]["worst_vs_rk4ref_in_tol_units"] = float(rel.max())
        print(name, json.dumps(out[name]), flush=True)
    # Path B: first 80 steps (anoxic -> aerobic switch at 51) + a terminal-like step is not timed here
    for name, mode, kw in (("B_dp45_1e-8", "dp45", dict(rtol=1e-8, atol=1e-10)), ("B_dp45_1e-6", "dp45", dict(rtol=1e-6, atol=1e-8)),
                           ("B_rk4_ref", "rk4", {}), ("B_rk4_sub20", "rk4", dict(rk4_sub_interval=20))):
        env = SbrOsVecEnv(n, device=dev, seed=2, mode=mode, **kw)
        t0 = time.perf_counter()
        env.reset()
        torch.cuda.synchronize()
        t_reset = time.perf_counter() - t0
        a = torch.stack([8 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen),
                         15 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen)], dim=1)
        per = []
        rhs = []
        for k in range(80):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            env.step_async(a)
            e1.record()
            torch.cuda.synchronize()
            per.append(e0.elapsed_time(e1))
            rhs.append(float(env.buf.counters[0].to(torch.float64).mean()))
        out[name] = dict(reset_ms=t_reset * 1e3, anoxic_ms=sum(per[5:45]) / 40, aerobic_ms=sum(per[55:80]) / 25,
                         switch_ms=per[51], rhs_anoxic=sum(rhs[5:45]) / 40, rhs_aerobic=sum(rhs[55:80]) / 25,
                         interval_steps_per_s_anoxic=n / (sum(per[5:45]) / 40) * 1e3,
                         interval_steps_per_s_aerobic=n / (sum(per[55:80]) / 25) * 1e3,
                         bad=int((env.buf.status != 0).sum()))
        print(name, json.dumps(out[name]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_modes.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
