"""Timing of the batch-to-batch (ILC) path: sbr_ilc_update and sbr_cycle_ilc per cycle at a few batch sizes (CUDA events)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200 import _abi, ilc
dev = torch.device("cuda:0")
for n in (4096, 1 << 15, 1 << 17):
    env = ilc.SbrIlcVecEnv(n, device=dev, seed=1, learn="feedback")
    env.reset()
    a = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5)) * 4 + 0.5
    env.step(a); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    # the two launches of a step, timed separately
    env._sp.copy_(a.t()); env._sp6.zero_(); env._sp6[2], env._sp6[4], env._sp6[5] = a.t()[0], a.t()[1], a.t()[2]
    ev[0].record()
    ilc.ilc_update(env.layout, env._w, env._D, env._sp6, env.so_learn, env.e_sum, env.e_last, env.u)
    ev[1].record()
    ilc.cycle_ilc(env.x, env.influent, env._sp, env.params, env.sched, env.layout, kla_base=env.kla_base, u=env.u, out=env._cyc)
    ev[2].record(); torch.cuda.synchronize()
    t_up, t_cy = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    for mode, name in ((_abi.MODE_RK4, "rk4"),):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ilc.cycle_ilc(env.x, env.influent, env._sp, env.params, env.sched, env.layout, kla_base=env.kla_base, u=env.u, out=env._cyc, mode=mode)
        e1.record(); torch.cuda.synchronize()
        t_rk4 = e0.elapsed_time(e1)
    S = int(env.layout.n_samples)
    rhs = float(env._cyc.counters[0].double().mean())
    print("n=%7d  S=%d  update %.3f ms (%.0f GB/s over 7 sample-row moves)  cycle dp45 %.2f ms (%.3g cycle-steps/s)  cycle rk4-grid %.2f ms  memory %.2f GB"
          % (n, S, t_up, 7 * S * n * 8 / t_up / 1e6, t_cy, n / t_cy * 1e3, t_rk4, 6 * S * n * 8 / 1e9), flush=True)
    del env
    torch.cuda.empty_cache()
