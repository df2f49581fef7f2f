"""Scratch A/B of DP45 controller variants: Path A cycle kernel (rtol 1e-7) and Path B step kernel (rtol 1e-8)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch, numpy as np
sys.path.insert(0, %r)
from gym_sbr2_b200 import _abi, core, schedule
from gym_sbr2_b200.vec_env import SbrV2VecEnv, SbrOsVecEnv
dev = torch.device("cuda:0")
res = {}
N = 1 << 20
env = SbrV2VecEnv(N, device=dev, seed=1)
env.reset()
a = torch.rand((N, 3), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
env.step_async(a); torch.cuda.synchronize()
ref = env._out.x_last.clone()
for rt, at in ((1e-6, 1e-8), (1e-7, 1e-9)):
    tol = _abi.make_tol(rt, at)
    core.cycle_v2(env.x0, env._loading, env._action, env.params, env.sched, out=env._out, mode=1, tol=tol); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = core.cycle_v2(env.x0, env._loading, env._action, env.params, env.sched, out=env._out, mode=1, tol=tol); e1.record(); torch.cuda.synchronize()
    cnt = o.counters.to(torch.float64)
    sc = torch.tensor([1.32, 30, 30, 1500, 150, 3000, 2000, 600, 8, 20, 20, 10, 10, 10], device=dev, dtype=torch.float64)[:, None]
    w = ((o.x_last - ref).abs() / (1e-5 * ref.abs() + 1e-9 * sc)).max(dim=0).values
    res["A_%%g" %% rt] = dict(ms=round(e0.elapsed_time(e1), 2), rhs=round(float(cnt[0].mean())), rej=round(float(cnt[1].mean()), 1),
                            p999=round(float(torch.quantile(w[:200000], 0.999)), 3), frac_gt1=float((w > 1).double().mean()))
del env
n = 1 << 20
env = SbrOsVecEnv(n, device=dev, seed=1, mode="dp45")
env.reset()
gen = torch.Generator(device=dev).manual_seed(1)
a = torch.stack([1 + 6 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen),
                 2 + 10 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen)], dim=1)
env._action.copy_(a.t())
per = []; rhs = []
for k in range(90):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); core.os_step(env.buf, env._action, env.params, env.sched, mode=env.mode, tol=env.tol); e1.record(); torch.cuda.synchronize()
    per.append(e0.elapsed_time(e1)); rhs.append(float(env.buf.counters[0].to(torch.float64).mean()))
an = sorted(per[5:45]); ae = sorted(per[55:90])
res["B"] = dict(anoxic_ms=round(an[len(an)//2], 4), aerobic_ms=round(ae[len(ae)//2], 4), rhs_an=round(sum(rhs[5:45])/40, 2), rhs_ae=round(sum(rhs[55:90])/35, 2))
print(json.dumps(res))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, SBR_B200_LIB=os.path.join(ROOT, lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-2000:], flush=True)
