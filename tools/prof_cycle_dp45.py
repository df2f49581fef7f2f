"""Profiling driver: the adaptive whole-cycle kernel on physically sorted inputs (what SbrV2VecEnv hands it), 2^20 envs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200.vec_env import SbrV2VecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
env = SbrV2VecEnv(n, device="cuda:0", seed=1, mode="dp45", rtol=1e-7, atol=1e-9)
env.reset()
a = torch.rand((n, 3), dtype=torch.float64, device="cuda:0", generator=torch.Generator(device="cuda:0").manual_seed(5))
for _ in range(2):
    o = env.step_async(a)
torch.cuda.synchronize()
print("ok", float(o.reward.mean()), float(o.counters[0].double().mean()))
