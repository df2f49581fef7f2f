"""Where an SBR-v4 episode's time goes: per-step CUDA-event times with and without the divergence-aware placement, the
steps that re-sort, and the host time per step (a loop that is host-bound shows GPU idle gaps)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from gym_sbr2_b200.vec_env import SbrV4VecEnv

n = 1 << 20
dev = "cuda:0"
for order in ("steps", "none", "steps"):
    env = SbrV4VecEnv(n, device=dev, seed=99, mode="dp45", order=order)
    gen = torch.Generator(device=dev).manual_seed(6)
    acts = [0.2 * torch.randn(n, dtype=torch.float64, device=dev, generator=gen) + 0.02 for _ in range(8)]
    env.reset()
    for k in range(3):
        env.step_async(acts[k])
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(495)]
    host = []
    ev[0].record()
    env.reset()
    t0 = time.perf_counter()
    for k in range(493):
        ev[k + 1].record()
        h0 = time.perf_counter()
        env.step_async(acts[k % 8])
        host.append(time.perf_counter() - h0)
    t_host = time.perf_counter() - t0
    ev[494].record()
    torch.cuda.synchronize()
    per = [ev[k + 1].elapsed_time(ev[k + 2]) for k in range(493)]
    big = [(k, round(p, 2)) for k, p in enumerate(per) if p > 1.0 and k > 26 and k < 492]
    hs = sorted(host)
    print("order=%s episode %.1f ms, sum steps %.1f, host loop issued in %.1f ms (median %.0f us, max %.1f ms per step), "
          "react median %.3f, steps > 1 ms in react: %s" % (order, ev[0].elapsed_time(ev[494]), sum(per), t_host * 1e3,
                                                            hs[len(hs) // 2] * 1e6, hs[-1] * 1e3,
                                                            sorted(per[40:490])[225], big[:30]), flush=True)
