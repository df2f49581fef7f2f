"""Which ordering of the envs gives the adaptive cycle kernel the fullest warps?  Runs the kernel once, then evaluates
candidate sort keys on the per-env RHS counts: lanes = 32 * mean(count) / mean over warps of max(count)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gym_sbr2_b200 import _abi, core
from gym_sbr2_b200.vec_env import SbrV2VecEnv
dev = torch.device("cuda:0")
N = 1 << 20
env = SbrV2VecEnv(N, device=dev, seed=1, mode="dp45", rtol=1e-7, atol=1e-9, order="none")
env.reset()
a = torch.rand((N, 3), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
o = env.step_async(a)
c = o.counters[0].to(torch.float64)
A = env._action
L = env._loading


def lanes(order):
    s = c[order]
    return round(32 * float(s.mean()) / float(s.view(-1, 32).max(dim=1).values.mean()), 2)


def two_level(primary, secondary, bins):
    b = torch.floor((primary - primary.min()) / (primary.max() - primary.min() + 1e-12) * bins).clamp_(0, bins - 1)
    s = (secondary - secondary.min()) / (secondary.max() - secondary.min() + 1e-12)
    return torch.argsort(b + 0.999 * s)


print("env order", lanes(torch.arange(N, device=dev)), "by count (ideal)", lanes(torch.argsort(c)))
print("sp3", lanes(torch.argsort(A[0])))
for name, sec in (("sp5", A[1]), ("sp8", A[2]), ("Ss_in", L[2]), ("Xs_in", L[4]), ("Snh_in", L[10]), ("Snd_in", L[11]),
                  ("cod_in", L[2] + L[4] + L[5]), ("n_in", L[10] + L[11] + L[12])):
    print(name, {bins: lanes(two_level(A[0], sec, bins)) for bins in (64, 256, 1024, 4096)})
# linear regression of the count on simple features within sp3 bins -> predicted count as the key
X = torch.stack([A[0], A[0] ** 2, 1 / (A[0] + 0.3), A[1], A[2], L[2], L[4], L[10], L[11], torch.ones_like(A[0])], dim=1)
w = torch.linalg.lstsq(X, c[:, None]).solution
pred = (X @ w)[:, 0]
print("linear model", lanes(torch.argsort(pred)), "corr", float(torch.corrcoef(torch.stack([pred, c]))[0, 1]))
for bins in (64, 256, 1024):
    b = torch.floor(A[0] / 8 * bins).clamp_(0, bins - 1).long()
    # per-bin regression on influent features
    feats = torch.stack([L[2], L[4], L[10], L[11], A[1], A[2], torch.ones_like(A[0])], dim=1)
    pred2 = torch.zeros_like(c)
    for k in range(bins):
        m = b == k
        if int(m.sum()) > 20:
            wk = torch.linalg.lstsq(feats[m], c[m][:, None]).solution
            pred2[m] = (feats[m] @ wk)[:, 0]
    key = b.double() * 1e6 + pred2
    print("per-bin regression", bins, lanes(torch.argsort(key)))
