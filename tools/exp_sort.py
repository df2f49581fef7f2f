"""Scratch experiment: does grouping similar envs into warps reduce DP45 divergence on the cycle kernel?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200 import _abi, core
from gym_sbr2_b200.vec_env import SbrV2VecEnv
dev = torch.device("cuda:0")
N = 1 << 20
env = SbrV2VecEnv(N, device=dev, seed=1)
env.reset()
a = torch.rand((N, 3), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
tol = _abi.make_tol(1e-7, 1e-9)
def run(action, loading, label):
    act = action.t().contiguous()
    out = core.CycleV2Out(N, dev)
    core.cycle_v2(env.x0, loading, act, env.params, env.sched, out=out, mode=1, tol=tol); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); core.cycle_v2(env.x0, loading, act, env.params, env.sched, out=out, mode=1, tol=tol); e1.record(); torch.cuda.synchronize()
    c = out.counters.to(torch.float64)[0]
    wm = c.view(-1, 32).max(dim=1).values.mean()
    print("%-28s %.2f ms  rhs mean %.0f  warp-max mean %.0f" % (label, e0.elapsed_time(e1), c.mean(), wm), flush=True)
load = env.influent.clone(); load[0] = env.fill_flow
run(a, load, "random order")
for key, name in ((a[:, 0], "sorted by action[0]"), (a[:, 0] * 1.0 + a[:, 1] * 0.2, "by a0 + 0.2 a1"), (load[10], "sorted by influent Snh")):
    idx = torch.argsort(key)
    run(a[idx], load[:, idx].contiguous(), name)
# two-level: bucket a0 into 1024 bins, then sort by a1 inside
idx = torch.argsort((a[:, 0] * 1024).floor() * 2 + a[:, 1])
run(a[idx], load[:, idx].contiguous(), "a0 bins (1024) then a1")
