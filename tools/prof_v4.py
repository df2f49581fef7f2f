"""Profiling driver for the SBR-v4 step kernel: reset + N env.steps at 2^20 envs (for ncu -k/-s/-c)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200.vec_env import SbrV4VecEnv
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 70
n = 1 << 20
env = SbrV4VecEnv(n, device="cuda:0", seed=1)
env.reset()
a = 0.05 * torch.randn(n, dtype=torch.float64, device="cuda:0") + 0.02
for k in range(steps):
    env.step_async(a)
torch.cuda.synchronize()
print("ok", float(env.buf.reward.mean()))
