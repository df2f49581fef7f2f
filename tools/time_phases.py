"""Scratch: per-launch times of whole SBROS-v1 / SBR-v4 episodes (reset, plain steps, phase switches, terminal step)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_sbr2_b200 import core
from gym_sbr2_b200.vec_env import SbrOsVecEnv, SbrV4VecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = "cuda:0"
out = {}


def ev():
    return torch.cuda.Event(enable_timing=True)


for mode, kw in (("dp45", {}), ("rk4", dict(rk4_sub_interval=20))):
    env = SbrOsVecEnv(n, device=dev, seed=77, mode=mode, **kw)
    gen = torch.Generator(device=dev).manual_seed(5)
    acts = [torch.stack([1 + 6 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen),
                         2 + 10 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen)], dim=0).contiguous()
            for _ in range(8)]
    infl = env._draw_influent()
    env.reset(influent=infl)
    for k in range(3):
        env.step_soa(acts[k])
    torch.cuda.synchronize()
    res = {}
    for rep in range(2):
        e = [ev() for _ in range(466)]
        e[0].record()
        env.reset(influent=infl)
        e[1].record()
        rhs_reset = env.buf.counters[0].to(torch.float64).mean().item()
        e[1].record()
        for k in range(463):
            env.step_soa(acts[k % 8])
            e[k + 2].record()
        torch.cuda.synchronize()
        per = [e[k + 1].elapsed_time(e[k + 2]) for k in range(463)]
        plain = sorted(per[60:270] + per[280:455])
        res = dict(ms_reset=e[0].elapsed_time(e[1]), rhs_reset=rhs_reset, ms_plain=plain[len(plain) // 2],
                   ms_first=per[0], ms_51=per[51], ms_52=per[52], ms_275=per[275], ms_276=per[276], ms_terminal=per[462],
                   ms_sum_steps=sum(per), rhs_terminal=env.buf.counters[0].to(torch.float64).mean().item(),
                   rhs_terminal_max=int(env.buf.counters[0].max()), bad=int((env.buf.status != 0).sum()))
    out["os_" + mode] = res
    del env

env = SbrV4VecEnv(n, device=dev, seed=3, mode="dp45")
gen = torch.Generator(device=dev).manual_seed(6)
acts = [(0.6 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen) - 0.25) for _ in range(8)]
for rep in range(2):
    e = [ev() for _ in range(496)]
    e[0].record()
    env.reset()
    e[1].record()
    rh = []
    for k in range(493):
        env.step_async(acts[k % 8])
        e[k + 2].record()
        if k in (5, 20, 30, 60, 100, 300, 492):
            c = env.buf.counters[0].to(torch.float64)
            rh.append((k, c.mean().item(), c.max().item()))
            e[k + 2].record()
    torch.cuda.synchronize()
    per = [e[k + 1].elapsed_time(e[k + 2]) for k in range(493)]
    out["v4_dp45"] = dict(ms_reset=e[0].elapsed_time(e[1]), ms_fill_median=sorted(per[:26])[13], ms_fill_sum=sum(per[:26]),
                          ms_react_median=sorted(per[26:492])[233], ms_react_sum=sum(per[26:492]),
                          ms_react_max=max(per[26:492]), ms_terminal=per[492], rhs=rh,
                          slowest=sorted(range(493), key=lambda k: -per[k])[:12],
                          per_first40=[round(p, 3) for p in per[:40]],
                          bad=int((env.buf.status != 0).sum()))
print(json.dumps(out))
