"""Scratch A/B: per-step time of the interval-per-step kernel for library variants (SBR_B200_LIB)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
from gym_sbr2_b200.vec_env import SbrOsVecEnv
n = 1 << 20
res = {}
for mode, kw in (("dp45", {}), ("rk4", dict(rk4_sub_interval=20)), ("rk4", dict(rk4_sub_interval=0))):
    env = SbrOsVecEnv(n, device="cuda:0", seed=1, mode=mode, **kw)
    env.reset()
    gen = torch.Generator(device="cuda:0").manual_seed(1)
    a = torch.stack([1 + 6 * torch.rand(n, dtype=torch.float64, device="cuda:0", generator=gen),
                     2 + 10 * torch.rand(n, dtype=torch.float64, device="cuda:0", generator=gen)], dim=1)
    env._action.copy_(a.t())
    from gym_sbr2_b200 import core
    per = []
    for k in range(90):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        core.os_step(env.buf, env._action, env.params, env.sched, mode=env.mode, tol=env.tol)
        e1.record(); torch.cuda.synchronize()
        per.append(e0.elapsed_time(e1))
    an = sorted(per[5:45]); ae = sorted(per[55:90])
    res["%%s_%%s" %% (mode, kw.get("rk4_sub_interval", ""))] = dict(anoxic_ms=an[len(an)//2], aerobic_ms=ae[len(ae)//2])
print(json.dumps(res))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, SBR_B200_LIB=os.path.join(ROOT, lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(lib, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-2000:], flush=True)
