"""Scratch A/B of library variants on the cycle kernel: time + RHS accuracy against the reference samples."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch, numpy as np
sys.path.insert(0, %r)
from gym_sbr2_b200 import _abi, core, schedule
from gym_sbr2_b200.vec_env import SbrV2VecEnv
dev = torch.device("cuda:0")
s = np.load(%r + "/tests/golden/stage_samples.npz")
p = _abi.default_params()
x = torch.as_tensor(np.ascontiguousarray(s["x"].T)).to(dev); n = x.shape[1]
load = torch.as_tensor(np.tile(s["load"][:, None], (1, n))).to(dev)
kla, ec = torch.as_tensor(s["kla"]).to(dev), torch.as_tensor(s["ec"]).to(dev)
res = {}
for tail, ref in ((0, s["d_react"]), (1, s["d_fill"]), (2, s["d_ec"])):
    dx = core.rhs(x, kla, p, tail, ec=ec, loading=load).cpu().numpy().T
    scale = np.abs(ref).max(axis=1, keepdims=True)
    res["rhs_err_tail%%d" %% tail] = float((np.abs(dx - ref) / (np.abs(ref) + 1e-3 * scale)).max())
N = 1 << 20
env = SbrV2VecEnv(N, device=dev, seed=1)
env.reset()
a = torch.rand((N, 3), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
env.step_async(a); torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    core.cycle_v2(env.x0, env._loading, env._action, env.params, env.sched, out=env._out, mode=0)
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
res["rk4_ms"] = min(ts)
res["reward_mean"] = float(env._out.reward.mean()); res["x_last_sum"] = float(env._out.x_last.sum())
import os
from gym_sbr2_b200 import parity
base = "/tmp/ab_xlast.pt"
if not os.path.exists(base):
    torch.save(dict(x=env._out.x_last.clone(), r=env._out.reward.clone()), base)
else:
    b = torch.load(base)
    scale = torch.as_tensor(parity.STATE_SCALE, device=dev, dtype=torch.float64)[:, None] if hasattr(parity, "STATE_SCALE") else torch.ones((14, 1), device=dev, dtype=torch.float64)
    u = (env._out.x_last - b["x"]).abs() / (1e-5 * b["x"].abs() + 1e-9 * scale)
    res["vs_first_units_max"] = float(u.max()); res["vs_first_units_p999"] = float(u.flatten().kthvalue(int(0.999 * u.numel())).values)
    res["vs_first_rel_max"] = float(((env._out.x_last - b["x"]).abs() / (b["x"].abs() + 1e-12)).max())
    res["reward_absdiff_max"] = float((env._out.reward - b["r"]).abs().max())
tol = _abi.make_tol(1e-7, 1e-9)
core.cycle_v2(env.x0, env._loading, env._action, env.params, env.sched, out=env._out, mode=1, tol=tol); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); core.cycle_v2(env.x0, env._loading, env._action, env.params, env.sched, out=env._out, mode=1, tol=tol); e1.record(); torch.cuda.synchronize()
res["dp45_1e-7_ms"] = e0.elapsed_time(e1)
print(json.dumps(res))
''' % (ROOT, ROOT)
for lib in sys.argv[1:]:
    env = dict(os.environ, SBR_B200_LIB=os.path.join(ROOT, lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(lib, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-2000:], flush=True)
