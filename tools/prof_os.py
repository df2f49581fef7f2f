"""Profiling driver for the interval-per-step kernels: reset + N env.steps at 2^20 envs (for ncu -k/-s/-c)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from gym_sbr2_b200.vec_env import SbrOsVecEnv

mode = sys.argv[1] if len(sys.argv) > 1 else "dp45"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 70
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
kw = dict(rk4_sub_interval=20) if mode == "rk4" else {}
if mode == "rk4grid":
    mode, kw = "rk4", {}
env = SbrOsVecEnv(n, device="cuda:0", seed=1, mode=mode, **kw)
env.reset()
gen = torch.Generator(device="cuda:0").manual_seed(1)
a = torch.stack([1 + 6 * torch.rand(n, dtype=torch.float64, device="cuda:0", generator=gen),
                 2 + 10 * torch.rand(n, dtype=torch.float64, device="cuda:0", generator=gen)], dim=1)
for k in range(steps):
    env.step_async(a)
torch.cuda.synchronize()
print("ok", float(env.buf.reward.mean()))
