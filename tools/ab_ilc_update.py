"""A/B of library variants (SBR_B200_LIB) on the batch-to-batch update kernel: time and GB/s at three batch sizes."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
from gym_sbr2_b200 import ilc
res = {}
for n in (4096, 1 << 15, 1 << 17):
    env = ilc.SbrIlcVecEnv(n, device="cuda:0", seed=1, learn="feedback")
    env.reset()
    a = torch.rand((n, 3), dtype=torch.float64, device="cuda:0") * 4 + 0.5
    env.step(a); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ilc.ilc_update(env.layout, env._w, env._D, env._sp6, env.so_learn, env.e_sum, env.e_last, env.u)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[2]
    res[str(n)] = dict(ms=round(t, 3), gbs=round(7 * 4769 * n * 8 / t / 1e6))
    u = env.u.clone()
    del env; torch.cuda.empty_cache()
res["u_checksum"] = float(u.double().abs().sum())
print(json.dumps(res))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, SBR_B200_LIB=os.path.join(ROOT, lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-1500:], flush=True)
