"""A/B of library variants on the adaptive (DP45) whole-cycle kernel: time, RHS / reject counters, lanes (warp max
vs mean of the per-env step count), x_last against the first variant and against the CPU twin (first 4096 envs).

    python tools/ab_cycle_dp45.py gym_sbr2_b200/libsbr_b200.so gym_sbr2_b200/_variants/libsbr_mb6.so ...
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, os, torch, numpy as np
sys.path.insert(0, %r)
from gym_sbr2_b200 import _abi, core, schedule
from gym_sbr2_b200.vec_env import SbrV2VecEnv
dev = torch.device("cuda:0")
N = int(os.environ.get("AB_N", str(1 << 20)))
env = SbrV2VecEnv(N, device=dev, seed=1)
env.reset()
a = torch.rand((N, 3), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
env.step_async(a); torch.cuda.synchronize()
res = {}
sc = torch.tensor([1.32, 30, 30, 1500, 150, 3000, 2000, 600, 8, 20, 20, 10, 10, 10], device=dev, dtype=torch.float64)[:, None]
perm = torch.argsort(env._action[0])
# physically sorted copies of the inputs (what SbrV2VecEnv.step_soa hands the kernel in adaptive mode)
sx0, sload, sact = env.x0[:, perm].contiguous(), env._loading[:, perm].contiguous(), env._action[:, perm].contiguous()
sout = core.CycleV2Out(N, dev)
for rt, at in ((1e-7, 1e-9), (1e-6, 1e-8)):
    tol = _abi.make_tol(rt, at)
    for name, pm in (("env", None), ("sorted", perm)):
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if pm is None:
                o = core.cycle_v2(env.x0, env._loading, env._action, env.params, env.sched, out=env._out, mode=1, tol=tol)
            else:
                o = core.cycle_v2(sx0, sload, sact, env.params, env.sched, out=sout, mode=1, tol=tol)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        cnt = o.counters.to(torch.float64)
        if pm is not None:
            env._out.x_last[:, perm] = sout.x_last
            cnt = torch.empty_like(cnt); cnt[:, perm] = o.counters.to(torch.float64)
        key = "%%g_%%s" %% (rt, name)
        res[key] = dict(ms=round(min(ts), 2), rhs=round(float(cnt[0].mean()), 1), rej=round(float(cnt[1].mean()), 1),
                        bad=int((o.status != 0).sum()))
        if pm is not None:
            srt = cnt[0][pm]
            res[key]["lanes_total_max"] = round(32 * float(srt.mean()) / float(srt.view(-1, 32).max(dim=1).values.mean()), 2)
    base = "/tmp/ab_dp45_xlast_%%g.pt" %% rt
    x = env._out.x_last.clone()
    if not os.path.exists(base):
        torch.save(x, base)
    else:
        b = torch.load(base)
        u = (x - b).abs() / (1e-5 * b.abs() + 1e-9 * sc)
        res["%%g_vs_first_units_max" %% rt] = float(u.max())
    if rt == 1e-7:
        from oracle.twin import binding as twin
        m = 4096
        r = twin.cycle_v2(env.x0[:, :m].cpu().numpy(), env._loading[:, :m].cpu().numpy(), env._action[:, :m].cpu().numpy(),
                          twin.default_params(), env.sched, mode=1, tol=tol)
        xt = torch.as_tensor(r["x_last"], device=dev)
        u = (x[:, :m] - xt).abs() / (1e-5 * xt.abs() + 1e-9 * sc)
        res["twin_units_max"] = float(u.max())
        res["twin_rhs_equal_frac"] = float((torch.as_tensor(r["counters"][0].astype(np.int64), device=dev) == cnt[0, :m].long()).double().mean())
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); core.cycle_v2(env.x0, env._loading, env._action, env.params, env.sched, out=env._out, mode=0); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
res["rk4_ms"] = round(min(ts), 2)
print(json.dumps(res))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, SBR_B200_LIB=os.path.join(ROOT, lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-3000:], flush=True)
