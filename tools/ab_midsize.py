"""Scratch A/B: RK4 cycle kernel time at mid-size batches for library variants."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
from gym_sbr2_b200 import core
from gym_sbr2_b200.vec_env import SbrV2VecEnv
dev = torch.device("cuda:0")
res = {}
for N in (4096, 32768, 65536, 75776, 131072, 262144, 1 << 20):
    env = SbrV2VecEnv(N, device=dev, seed=1)
    env.reset()
    a = torch.rand((N, 3), dtype=torch.float64, device=dev)
    env.step_async(a); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); core.cycle_v2(env.x0, env._loading, env._action, env.params, env.sched, out=env._out, mode=0); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[N] = round(min(ts), 3)
print(json.dumps(res))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, SBR_B200_LIB=os.path.join(ROOT, lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-1500:], flush=True)
