"""How much faster is sbr_v4_step when the envs of a warp need similar numbers of steps?  At a few points of an episode:
time the step in env order, then physically reorder every per-env buffer by the previous step's RHS count and time the
same step again (upper bound of what keeping the state sorted can recover; I/O stays unit-stride here)."""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200 import _abi, core
from gym_sbr2_b200.vec_env import SbrV4VecEnv

n = 1 << 20
dev = "cuda:0"
env = SbrV4VecEnv(n, device=dev, seed=3, mode="dp45")
gen = torch.Generator(device=dev).manual_seed(6)
acts = [0.2 * torch.randn(n, dtype=torch.float64, device=dev, generator=gen) + 0.02 for _ in range(8)]
env.reset()


def timed(buf, loading, action):
    st0, done0 = buf.st.clone(), buf.done.clone()
    ts = []
    for _ in range(3):
        buf.st.copy_(st0); buf.done.copy_(done0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        core.v4_step(buf, loading, action, env.params, env.sched, mode=env.mode, tol=env.tol)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    buf.st.copy_(st0); buf.done.copy_(done0)
    return min(ts)


for k in range(493):
    if k in (5, 15, 25, 60, 120, 200, 300, 400):
        a = acts[k % 8].contiguous()
        prev = env.buf.counters[0].clone()
        t_env = timed(env.buf, env._loading, a)
        perm = torch.argsort(prev)
        sb = core.V4Buffers(n, dev)
        sb.st.copy_(env.buf.st[:, perm]); sb.done.copy_(env.buf.done[perm])
        t_sorted = timed(sb, env._loading[:, perm].contiguous(), a[perm].contiguous())
        print(dict(step=k, ms_env_order=round(t_env, 3), ms_sorted_by_prev_count=round(t_sorted, 3)), flush=True)
    env.step_async(acts[k % 8])
