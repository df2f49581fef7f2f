import sys, time, torch
sys.path.insert(0, ".")
from gym_sbr2_b200 import ilc
dev = torch.device("cuda:0")
for learn in ("feedback", "frozen"):
    n = 4096
    env = ilc.SbrIlcVecEnv(n, device=dev, seed=2, learn=learn)
    env.reset()
    g = torch.Generator(device=dev).manual_seed(1)
    t0 = time.time()
    for c in range(20):
        a = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=g) * 6 - 0.5     # outside the box too
        a[::97, 0] = 0.0                                                                   # exact zeros: the 0/0 quirk
        obs, r, d, info = env.step(a)
    torch.cuda.synchronize()
    st = info["status"]
    print(learn, "20 cycles %.2f s; envs flagged non-finite %d, step-limit %d; finite obs rows %d of %d; rhs/env %.0f"
          % (time.time() - t0, int((st & 1).bool().sum()), int((st & 4).bool().sum()), int(torch.isfinite(obs).all(dim=1).sum()), n,
             float(info["counters"][0].double().mean())), flush=True)
# the same plant under the feedback PID alone (SBR-v1), 50 chained cycles
env = ilc.SbrV1VecEnv(4096, device=dev, seed=2)
env.reset()
g = torch.Generator(device=dev).manual_seed(1)
t0 = time.time()
for c in range(50):
    a = torch.rand((4096, 3), dtype=torch.float64, device=dev, generator=g) * 6 - 0.5
    obs, r, d, info = env.step(a)
torch.cuda.synchronize()
st = info["status"]
x = info["x_last"]
print("SBR-v1 50 cycles %.2f s; non-finite %d, step-limit %d; V in [%.4f, %.4f], Xbh in [%.0f, %.0f], rhs/env %.0f"
      % (time.time() - t0, int((st & 1).bool().sum()), int((st & 4).bool().sum()), float(x[0].min()), float(x[0].max()),
         float(x[5].min()), float(x[5].max()), float(info["counters"][0].double().mean())), flush=True)
