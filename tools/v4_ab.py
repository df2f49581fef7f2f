"""Scratch A/B: SBR-v4 per-step kernel times for library variants (SBR_B200_LIB)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
from gym_sbr2_b200.vec_env import SbrV4VecEnv
n = 1 << 20; dev = "cuda:0"
import os
env = SbrV4VecEnv(n, device=dev, seed=3, mode="dp45", order=os.environ.get("V4_ORDER", "auto"))
env.place_group = int(os.environ.get("V4_GROUP", "4"))
gen = torch.Generator(device=dev).manual_seed(6)
if os.environ.get("V4_ACTS", "bench") == "bench":
    acts = [0.2 * torch.randn(n, dtype=torch.float64, device=dev, generator=gen) + 0.02 for _ in range(8)]
else:
    acts = [(0.6 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen) - 0.25) for _ in range(8)]
for rep in range(2):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(496)]
    env.reset()
    e[1].record()
    for k in range(493):
        env.step_async(acts[k %% 8]); e[k + 2].record()
    torch.cuda.synchronize()
    per = [e[k + 1].elapsed_time(e[k + 2]) for k in range(493)]
print(json.dumps(dict(order=env.order, group=env.place_group, fill_sum=sum(per[:26]), react_median=sorted(per[26:492])[233], react_sum=sum(per[26:492]), terminal=per[492], total=sum(per))))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, SBR_B200_LIB=os.path.join(ROOT, lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(lib, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-2000:], flush=True)
