"""SBR-v4 episode at 2^20 envs behind the policy head, three ways: step by step with env-indexed buffers (collect_episode_v4),
step by step with every buffer in slot order (collect_episode_v4_sorted, re-sort every R steps), fused (K = 8)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200 import rollout
from gym_sbr2_b200.vec_env import SbrV4VecEnv
dev = torch.device("cuda:0")
n = 1 << 20
env = SbrV4VecEnv(n, device=dev, seed=1)
pol = rollout.TinyPolicy(dev, n_in=14, lo=(-0.1,), span=(0.4,), seed=3)
def timed(fn, *a, **k):
    fn(*a, **k); torch.cuda.synchronize()           # warm-up episode (first re-sort pays one-time set-up)
    env.epoch.zero_()
    t0 = time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, r
env.epoch.zero_()
t_a, ra = timed(rollout.collect_episode_v4, env, pol)
print("step by step, env-indexed buffers (slot placement of the state): %.1f ms" % t_a, flush=True)
for R in (4, 8, 16, 32):
    env.epoch.zero_()
    t_b, rb = timed(rollout.collect_episode_v4_sorted, env, pol, resort_every=R)
    same = bool(torch.equal(ra["returns"], rb["returns"]))
    print("step by step, all buffers in slot order, re-sort every %2d steps: %.1f ms  (returns bit-identical: %s)" % (R, t_b, same), flush=True)
env.epoch.zero_()
t_c, rc = timed(rollout.collect_episode_v4_fused, env, pol, K=8)
print("fused rollout K = 8: %.1f ms  (max |d returns| %.2e)" % (t_c, float((rc["returns"] - ra["returns"]).abs().max())))
