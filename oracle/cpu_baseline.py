"""CPU baseline = the oracle port of the reference's scipy-odeint path, timed on the host cores.
TEST / MEASUREMENT INFRASTRUCTURE ONLY: imported by bench.py's `cpu_baseline` leg and `--impl reference` arm.

One Python process per core (BLAS/OpenMP threads pinned to 1), each stepping independent seeded SBR-v2 episodes
(reset influent draw + one whole-cycle step) exactly as N reference processes would.
The unmodified reference cannot travel to the GPU box (/root/reference is absent there), so kind = "port": the
port reproduces the reference bit for bit on this path (tests/test_oracle_golden.py) and has the same cost
structure (530 odeint calls, ~13.6k Python RHS callbacks per cycle).
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
