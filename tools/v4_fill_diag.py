"""RHS evaluations per env and step over the first 40 steps of an SBR-v4 episode (26 fill steps, then react)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200.vec_env import SbrV4VecEnv
n = 1 << 18
env = SbrV4VecEnv(n, device="cuda:0", seed=99, order="none")
env.reset()
gen = torch.Generator(device="cuda:0").manual_seed(6)
for k in range(40):
    a = 0.2 * torch.randn(n, dtype=torch.float64, device="cuda:0", generator=gen) + 0.02
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step_async(a); e1.record(); torch.cuda.synchronize()
    c = env.buf.counters[0].double()
    w = c.view(-1, 32).max(dim=1).values
    print("step %2d  %.3f ms  rhs mean %.1f  p50 %.0f  p99 %.0f  max %.0f  warp-max mean %.1f  rejects %.2f  u mean %.3f" % (
        k, e0.elapsed_time(e1), float(c.mean()), float(c.median()), float(torch.quantile(c, 0.99)), float(c.max()),
        float(w.mean()), float(env.buf.counters[1].double().mean()), float(env.buf.st[15].mean())))
