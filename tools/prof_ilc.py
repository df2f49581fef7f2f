"""Profiling driver for the batch-to-batch (SBR-v0) kernels: reset (cycle 0) + 2 steps at 2^16 envs (for ncu -k / -s / -c)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200 import ilc
n = 1 << 16
env = ilc.SbrIlcVecEnv(n, device="cuda:0", seed=1, learn="feedback")
env.reset()
a = torch.rand((n, 3), dtype=torch.float64, device="cuda:0", generator=torch.Generator(device="cuda:0").manual_seed(5)) * 4 + 0.5
for _ in range(2):
    o, r, d, info = env.step(a)
torch.cuda.synchronize()
print("ok", float(r.mean()), float(env._cyc.counters[0].double().mean()))
