"""compute-sanitizer driver: every entry point once on ragged batch sizes (run as
`compute-sanitizer --tool memcheck python tools/sanitize.py`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200 import _abi, core
from gym_sbr2_b200.vec_env import SbrV2VecEnv, SbrOsVecEnv, SbrV4VecEnv

dev = torch.device("cuda:0")
for n in (1, 77, 333):
    for mode in ("rk4", "dp45"):
        e = SbrV2VecEnv(n, device=dev, seed=1, mode=mode)
        e.reset()
        e.step(torch.rand((n, 3), dtype=torch.float64, device=dev))
        o = SbrOsVecEnv(n, device=dev, seed=2, mode=mode)
        o.reset()
        for k in range(3):
            o.step(torch.rand((n, 2), dtype=torch.float64, device=dev) * 5)
        o.reset(mask=(torch.arange(n, device=dev) % 2 == 0))
        v = SbrV4VecEnv(n, device=dev, seed=3, mode=mode)
        v.reset()
        for k in range(3):
            v.step(torch.rand((n, 1), dtype=torch.float64, device=dev) * 0.1)
    x = e.x0.clone()
    kla = torch.rand(n, dtype=torch.float64, device=dev) * 100
    core.rhs(x, kla, e.params, _abi.TAIL_FILL, loading=e._loading)
    core.integrate_interval(x, kla, e.params, _abi.TAIL_EC, 0.02 / 24, 10, ec=kla * 1e-6)
    core.reward_stats(e._out.reward, e._out.status)
torch.cuda.synchronize()
print("sanitize driver done")
