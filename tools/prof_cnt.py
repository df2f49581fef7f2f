"""Profiling driver for the SBRCnt / SBROS-v2 step kernel: reset + N env.steps at 2^20 envs (for ncu -k/-s/-c)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from gym_sbr2_b200.cnt import SbrCntVecEnv

kind = sys.argv[1] if len(sys.argv) > 1 else "ma1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
env = SbrCntVecEnv(kind, n, device="cuda:0", seed=1)
env.reset()
g = torch.Generator(device="cuda:0").manual_seed(1)
per = []
for k in range(steps):
    if kind == "os2":
        a = torch.stack([1 + 2 * torch.rand(n, dtype=torch.float64, device="cuda:0", generator=g),
                         torch.zeros(n, dtype=torch.float64, device="cuda:0")], dim=1)
    else:
        # per-env DO set-points ramping to U(1, 3) g/m3 over the first aerobic steps, small random moves afterwards
        up = range(60, 64) if kind == "ma1" else range(1, 5)
        a = 0.01 * torch.randn(n, dtype=torch.float64, device="cuda:0", generator=g) * (0.1 if kind == "cnt0" else 1.0)
        if k in up:
            a = a + (0.25 + 0.5 * torch.rand(n, dtype=torch.float64, device="cuda:0", generator=g)) * (0.06 if kind == "cnt0" else 1.0)
        if kind == "ma1" and k < 60 or kind == "ma1" and k > 300:
            a = torch.zeros(n, dtype=torch.float64, device="cuda:0")
        if k == 0 and kind in ("cnt2", "ma1"):
            a = torch.full((n,), -2.0, dtype=torch.float64, device="cuda:0")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.step_async(a)
    e1.record()
    torch.cuda.synchronize()
    per.append(e0.elapsed_time(e1))
tail = sorted(per[(2 * len(per)) // 3:])
print("ok %s median step %.4f ms (last third of %d steps), first %.3f ms" % (kind, tail[len(tail) // 2], steps, per[0]))
