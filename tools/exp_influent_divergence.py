"""How much of the adaptive cycle kernel's residual warp divergence comes from the per-env INFLUENT (which the divergence-aware
order ignores)?  Same action for every env, per-env influent draws: lanes = 32 * mean(RHS) / mean over warps of max(RHS)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200.vec_env import SbrV2VecEnv
dev = torch.device("cuda:0")
n = 1 << 17
env = SbrV2VecEnv(n, device=dev, seed=1, mode="dp45", rtol=1e-7, atol=1e-9, order="none")
env.reset()
for act in ([0.05, 0.5, 0.5], [0.3, 0.5, 0.5], [0.8, 0.2, 0.9]):
    a = torch.tensor([act] * n, dtype=torch.float64, device=dev)
    o = env.step_async(a); torch.cuda.synchronize()
    c = o.counters[0].double()
    lanes = 32 * float(c.mean()) / float(c.view(-1, 32).max(dim=1).values.mean())
    print("action %s: RHS mean %.0f  std %.0f  min %.0f  max %.0f  -> %.2f of 32 lanes from influent alone"
          % (act, c.mean(), c.std(), c.min(), c.max(), lanes), flush=True)
