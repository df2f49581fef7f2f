"""CPU baseline, timed on the host cores.  TEST / MEASUREMENT INFRASTRUCTURE ONLY: imported by bench.py's
`cpu_baseline` leg and `--impl reference` arm.

kind = "reference": the UNMODIFIED reference package, installed by oracle/install_ref.py into baseline/_ref/
(git-ignored, travels to the GPU box with the snapshot) and imported through oracle/ref_shim.py (only `gym` and
`matplotlib` are stubbed; numpy and scipy.integrate.odeint are the real ones; the envs' per-step prints are swallowed).
One Python process per core (BLAS/OpenMP threads pinned to 1), each running independent seeded episodes --
`np.random.seed(s)`, `SbrEnv2().reset()`, `step(action)` -- exactly as N reference processes would (BASELINE.md
section 3).  kind = "port" (`run`, `run_os`): the oracle restatement (oracle/sbr_oracle.py), which reproduces the
reference bit for bit on this path (tests/test_oracle_golden.py) but skips its trajectory appends and prints; it is
reported beside the reference's number and is the fallback when baseline/_ref is missing.
"""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(args):
    seed, n_steps = args
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import numpy as np
    from gym_sbr2_b200 import influent
    from oracle import sbr_oracle as O
    rng = np.random.RandomState(seed)
    t0 = time.perf_counter()
    acc = 0.0
    for _ in range(n_steps):
        infl = influent.sample_numpy(0, rng)
        try:
            out = O.sbr_v2_step(rng.rand(3), infl)
            acc += out["reward"]
        except (OverflowError, ValueError, ZeroDivisionError, FloatingPointError):
            pass        # the reference's own failure regimes: the step was paid for, count it
    return time.perf_counter() - t0, acc


def reference_root():
    """Where the unmodified reference can be imported from: baseline/_ref (installed copy), else the authoring
    container's read-only checkout, else None."""
    for root in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(root, "gym_SBR", "envs", "gym_SBR_env2.py")):
            return root
    return None


def _load_reference(root):
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    os.environ["SBR_REFERENCE_ROOT"] = root
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_shim
    ref_shim.REFERENCE_ROOT = root
    ref_shim.load_reference()
    return ref_shim


def _worker_ref(args):
    """SBR-v2 through the reference's own class: reset (influent draw) + one whole-cycle step per episode."""
    seed, n_steps, root = args
    ref_shim = _load_reference(root)
    import numpy as np
    from gym_SBR.envs.gym_SBR_env2 import SbrEnv2
    acc = 0.0
    with ref_shim.quiet():
        env = SbrEnv2()
        np.random.seed(seed)
        t0 = time.perf_counter()
        for _ in range(n_steps):
            try:
                env.reset()
                _, reward, _, _ = env.step(np.random.rand(3))
                acc += reward
            except (OverflowError, ValueError, ZeroDivisionError, FloatingPointError, NameError):
                pass    # the reference's own failure regimes: the step was paid for, count it
        dt = time.perf_counter() - t0
    return dt, acc


def _worker_ref_os(args):
    """SBROS-v1 through the reference's own class: one 72-s PID interval per step, 463 steps per episode."""
    seed, n_steps, root = args
    ref_shim = _load_reference(root)
    import warnings
    import numpy as np
    from gym_SBR.envs.gym_SBR_oneshot import SbrOS
    acc = 0.0
    with ref_shim.quiet(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = SbrOS()
        np.random.seed(seed)
        env.reset()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            try:
                out = env.step([2.0 + np.random.rand(), 4.0 + 2 * np.random.rand()])
                acc += out[2]
                done = out[3]
            except (OverflowError, ValueError, ZeroDivisionError, FloatingPointError, NameError):
                done = True     # e.g. round(inf) in the draw (gym_SBR_oneshot.py:2338): an RL loop would reset
            if done:
                env.reset()
        dt = time.perf_counter() - t0
    return dt, acc


def _run_pool(worker, jobs_warm, jobs, procs):
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        if jobs_warm:
            pool.map(worker, jobs_warm)
        t0 = time.perf_counter()
        res = pool.map(worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    return res, wall


def run_reference(steps_per_proc=3, procs=None, warmup=1, path="v2"):
    """The unmodified reference on every host core.  path "v2": SBR-v2 cycle-steps/s; "os": SBROS-v1
    interval-steps/s.  Returns None when no reference install is available."""
    root = reference_root()
    if root is None:
        return None
    procs = procs or usable_cores()
    worker = _worker_ref if path == "v2" else _worker_ref_os
    warm = [(10_000 + i, warmup, root) for i in range(procs)] if warmup else None
    res, wall = _run_pool(worker, warm, [(i, steps_per_proc, root) for i in range(procs)], procs)
    total = procs * steps_per_proc
    return dict(value=total / wall, cores=procs, steps=total, wall_s=wall, per_core=total / sum(r[0] for r in res),
                root=os.path.relpath(root, ROOT) if root.startswith(ROOT) else root)


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run(steps_per_proc=3, procs=None, warmup=1):
    """Returns dict(value cycle-steps/s, cores, steps, wall_s).  Wall time is that of the timed map only
    (workers are started and warmed first so imports are not counted)."""
    procs = procs or usable_cores()
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        if warmup:
            pool.map(_worker, [(10_000 + i, warmup) for i in range(procs)])
        t0 = time.perf_counter()
        res = pool.map(_worker, [(i, steps_per_proc) for i in range(procs)], chunksize=1)
        wall = time.perf_counter() - t0
    total = procs * steps_per_proc
    return dict(value=total / wall, cores=procs, steps=total, wall_s=wall,
                per_core=total / sum(r[0] for r in res))


def _worker_os(args):
    seed, n_steps = args
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import warnings
    import numpy as np
    from gym_sbr2_b200 import influent
    from oracle import sbr_oracle as O
    rng = np.random.RandomState(seed)
    env = O.SbrOsOracle()
    env.reset(influent.sample_numpy(6, rng))
    t0 = time.perf_counter()
    acc = 0.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in range(n_steps):
            try:
                (_, _), _, r, done = env.step([2.0 + rng.rand(), 4.0 + 2 * rng.rand()])
                acc += r
            except (OverflowError, ValueError, ZeroDivisionError, FloatingPointError):
                # the reference's own failure regime (round(inf) in the draw, gym_SBR_oneshot.py:2338): an RL loop
                # around the reference would catch it and start a new episode; the step was paid for, count it
                done = True
            if done:
                env.reset(influent.sample_numpy(6, rng))
    return time.perf_counter() - t0, acc


def run_os(steps_per_proc=400, procs=None):
    """SBROS-v1 interval-steps/s of the oracle port on the host cores (one env per process, 72-s PID intervals)."""
    procs = procs or usable_cores()
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_worker_os, [(20_000 + i, 20) for i in range(procs)])
        t0 = time.perf_counter()
        res = pool.map(_worker_os, [(i, steps_per_proc) for i in range(procs)], chunksize=1)
        wall = time.perf_counter() - t0
    total = procs * steps_per_proc
    return dict(value=total / wall, cores=procs, steps=total, wall_s=wall, per_core=total / sum(r[0] for r in res))
