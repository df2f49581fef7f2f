"""Scratch: per-step kernel time and RHS statistics over one SBROS-v1 episode."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_sbr2_b200.vec_env import SbrOsVecEnv
n = 1 << 20
dev = "cuda:0"
mode = sys.argv[1] if len(sys.argv) > 1 else "dp45"
kw = dict(rtol=float(sys.argv[2]), atol=float(sys.argv[3])) if len(sys.argv) > 3 else {}
env = SbrOsVecEnv(n, device=dev, seed=77, mode=mode, **kw)
gen = torch.Generator(device=dev).manual_seed(5)
acts = [torch.stack([1 + 6 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen),
                     2 + 10 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen)], dim=0).contiguous()
        for _ in range(8)]
infl = env._draw_influent()
env.reset(influent=infl)
for k in range(3):
    env.step_soa(acts[k])
env.reset(influent=infl)
torch.cuda.synchronize()
rows = []
for k in range(463):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.step_soa(acts[k % 8])
    e1.record()
    c = env.buf.counters[0].to(torch.float64)
    w = c.view(-1, 32).max(dim=1).values
    torch.cuda.synchronize()
    rows.append((k, round(e0.elapsed_time(e1), 3), round(c.mean().item(), 2), int(c.max()), round(w.mean().item(), 2),
                 round(env.buf.counters[1].to(torch.float64).mean().item(), 3)))
print("k ms rhs_mean rhs_max warpmax_mean rej_mean")
for r in rows:
    print(*r)
print("sum_ms", sum(r[1] for r in rows))
