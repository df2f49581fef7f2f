"""SBR-v4 step kernel: what a warp pays for per-env adaptive steps, and how much of it an ordering of the envs would
recover.  Per sampled step: mean RHS per env, mean over warps of the warp's max (env order), and the same if the envs
were sorted by the previous step's RHS count / by the set-point u / by the influent scenario."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200 import _abi
from gym_sbr2_b200.vec_env import SbrV4VecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
dev = "cuda:0"
env = SbrV4VecEnv(n, device=dev, seed=3, mode="dp45")
gen = torch.Generator(device=dev).manual_seed(6)
acts = [0.2 * torch.randn(n, dtype=torch.float64, device=dev, generator=gen) + 0.02 for _ in range(8)]
env.reset()
prev = None


def warp_max(c, order=None):
    c = c if order is None else c[order]
    return float(c.view(-1, 32).max(dim=1).values.mean())


for k in range(493):
    env.step_async(acts[k % 8])
    c = env.buf.counters[0].to(torch.float64)
    if k in (3, 10, 20, 30, 40, 60, 120, 200, 300, 400, 480) and prev is not None:
        u = env.buf.st[_abi.V4_U]
        so = env.buf.st[8]
        line = dict(step=k, mean=round(float(c.mean()), 2), max=float(c.max()), warp_env=round(warp_max(c), 2),
                    warp_by_prev=round(warp_max(c, torch.argsort(prev)), 2),
                    warp_by_u=round(warp_max(c, torch.argsort(u)), 2),
                    warp_by_scn=round(warp_max(c, torch.argsort(env.scenario.to(torch.int64) * 1000 + (u * 100).to(torch.int64))), 2),
                    warp_by_self=round(warp_max(c, torch.argsort(c)), 2),
                    corr_prev=round(float(torch.corrcoef(torch.stack([c, prev]))[0, 1]), 3))
        print(line, flush=True)
    prev = c.clone()
