"""Profiling driver for the fused rollout kernel (sbr_os_rollout_k): reset + L launches of K env.steps at 2^20 envs with
the policy head in-kernel (for ncu -k regex:sbr_os_step -s/-c)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from gym_sbr2_b200 import core, rollout
from gym_sbr2_b200.vec_env import SbrOsVecEnv

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 14
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
env = SbrOsVecEnv(n, device="cuda:0", seed=1, mode="dp45", emit=("obs_do", "obs_ec"))
env.reset()
policy = rollout.TinyPolicy("cuda:0")
pol = policy.as_struct()
b = env.buf
policy.act_into(b.obs_do, b.obs_ec, env._action)
rewards = torch.zeros((K, n), dtype=torch.float64, device="cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for j in range(launches):
    if j == launches - 1:
        e0.record()
    core.os_rollout_k(b, env._action, pol, rewards, env.params, env.sched, mode=env.mode, tol=env.tol, emit=env.emit)
e1.record()
torch.cuda.synchronize()
print("ok K=%d last launch %.3f ms (%.4f ms per env.step)" % (K, e0.elapsed_time(e1), e0.elapsed_time(e1) / K),
      float(rewards.mean()))
