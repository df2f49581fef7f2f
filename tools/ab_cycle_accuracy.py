"""Accuracy of library variants of the adaptive whole-cycle kernel on 2^20 random envs: x_last against RK4 with 40
sub-steps per PID interval (from the first library), in units of the parity tolerance (1e-5 relative + 1e-9 x_1_state).

    python tools/ab_cycle_accuracy.py gym_sbr2_b200/libsbr_b200.so gym_sbr2_b200/_variants/libsbr_x.so ...
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, os, torch
sys.path.insert(0, %r)
from gym_sbr2_b200 import _abi, core, schedule
from gym_sbr2_b200.vec_env import SbrV2VecEnv
dev = torch.device("cuda:0")
N = 1 << 20
env = SbrV2VecEnv(N, device=dev, seed=1, mode="dp45", rtol=1e-7, atol=1e-9)
env.reset()
a = torch.rand((N, 3), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
ref_path = "/tmp/ab_acc_ref.pt"
if not os.path.exists(ref_path):
    env._action.copy_(a.t())
    fine = core.cycle_v2(env.x0, env._loading, env._action, env.params, schedule.cycle_schedule(substeps=40), mode=_abi.MODE_RK4)
    torch.save(fine.x_last.clone(), ref_path)
ref = torch.load(ref_path)
scale = torch.tensor([1.32, 30, 30, 1500, 150, 3000, 2000, 600, 8, 20, 20, 10, 10, 10], dtype=torch.float64, device=dev)[:, None]
res = {}
for rt, at in ((1e-7, 1e-9), (1e-6, 1e-8)):
    env.tol = _abi.make_tol(rt, at)
    o = env.step_async(a); torch.cuda.synchronize()
    w = ((o.x_last - ref).abs() / (1e-5 * ref.abs() + 1e-9 * scale)).max(dim=0).values
    q = torch.quantile(w[:1 << 18], torch.tensor([0.5, 0.99, 0.999, 0.9999], dtype=torch.float64, device=dev))
    res["%%g" %% rt] = dict(median=round(float(q[0]), 5), p99=round(float(q[1]), 4), p999=round(float(q[2]), 4), p9999=round(float(q[3]), 3),
                         max=round(float(w.max()), 2), n_above_1=int((w > 1).sum()), n_above_01=int((w > 0.1).sum()),
                         rhs=round(float(o.counters[0].double().mean()), 1), bad=int((o.status != 0).sum()))
print(json.dumps(res))
''' % ROOT
if os.path.exists("/tmp/ab_acc_ref.pt"):
    os.remove("/tmp/ab_acc_ref.pt")
for lib in sys.argv[1:]:
    env = dict(os.environ, SBR_B200_LIB=os.path.join(ROOT, lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-2000:], flush=True)
