"""Scratch: launch-bound regime (BASELINE config[1] size, 4096 envs): per-step wall time through the Python API,
with and without a CUDA graph around policy + step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200 import core, rollout
from gym_sbr2_b200.vec_env import SbrOsVecEnv, SbrV2VecEnv

dev = torch.device("cuda:0")
for n in (4096, 65536):
    env = SbrOsVecEnv(n, device=dev, seed=1, mode="dp45")
    pol = rollout.TinyPolicy(dev)
    env.reset()
    a = torch.rand((n, 2), dtype=torch.float64, device=dev) * 4 + 1
    for _ in range(5):
        env.step(a)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        env.step(a)
    torch.cuda.synchronize(); t_api = (time.perf_counter() - t0) / 200
    env.reset()
    b = env.buf
    for _ in range(5):
        env.step_soa(pol.act_into(b.obs_do, b.obs_ec, env._action))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        env.step_soa(pol.act_into(b.obs_do, b.obs_ec, env._action))
    torch.cuda.synchronize(); t_pol = (time.perf_counter() - t0) / 200
    # CUDA graph: policy + step, 8 steps per replay
    env.reset()
    act = torch.zeros((2, n), dtype=torch.float64, device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            pol.act_into(b.obs_do, b.obs_ec, act); env.step_soa(act)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(8):
            pol.act_into(b.obs_do, b.obs_ec, act); env.step_soa(act)
    env.reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(25):
        g.replay()
    torch.cuda.synchronize(); t_graph = (time.perf_counter() - t0) / 200
    print("n=%d  step(a) %.1f us  policy+step_soa %.1f us  graph(policy+step) %.1f us per env.step  -> %.2f M interval-steps/s with graph"
          % (n, t_api * 1e6, t_pol * 1e6, t_graph * 1e6, n / t_graph / 1e6), flush=True)
    print("   steps done", float(b.st[33].max()), "status", int(b.status.max()))
env = SbrV2VecEnv(4096, device=dev, seed=1)
env.reset()
a3 = torch.rand((4096, 3), dtype=torch.float64, device=dev)
env.step(a3); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    env.step(a3)
torch.cuda.synchronize(); print("SBR-v2 n=4096: %.2f ms per cycle-step launch" % ((time.perf_counter() - t0) / 5 * 1e3))
