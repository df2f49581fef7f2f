#!/usr/bin/env python
"""bench.py -- SBR env-steps/sec on N B200s, FP64 roofline fraction, next to the CPU scipy-odeint path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--envs-per-gpu M] [--mode rk4|dp45]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...)

A "step" is one pass of the hot path over one batch: every env of the rank's shard advances by ONE SBR-v2
env-step = one whole 12-h cycle = 528 PID intervals + settle/draw (gym_SBR_env2.py:131-171), in ONE kernel launch.
Workload (BASELINE.json configs[4] at one GPU, named in config.workload): 2^20 envs per GPU, random DO set-point
actions, per-env influent draws, one cycle each; envs are independent, so ranks shard them with no data-path
collective ("scaling": "weak"); the only collective is an all_gather of 5 reward statistics per step.
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the flop/byte accounting.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sbr_env_steps_per_sec"
UNIT = "cycle-steps/s"          # one env-step of SBR-v2 = one 12-h cycle = 528 PID intervals (72 s each)
INTERVALS_PER_CYCLE = 528

# algorithmic flops (add/mul/div = 1, FMA = 2; CSE-minimal RHS)
F_REACT, F_FILL = 66, 106
OVH_REACT = 2 * (3 * 9 + 4 * 11)     # RK4 step: 3 stage-input + 4 accumulate FMAs over 9 / 11 active components
OVH_FILL = 2 * (3 * 14 + 4 * 14)
F_EPILOGUE = 400
# Dormand-Prince step on the 9 active components: 15 stage-input FMAs + 5 solution FMAs + 6 error FMAs per component,
# ~5 flops per component for the norm (SURVEY.md 8d counts these over all 14 components: 828 instead of 513)
DP45_STEP_OVH = 2 * 9 * 15 + 2 * 9 * 5 + 2 * 9 * 6 + 5 * 9
BYTES_PER_ENV = (14 + 14 + 3) * 8 + (14 + 3 + 1 + 12) * 8 + 4 + 8   # SoA reads + writes per env per cycle


def cycle_flops_rk4(sched):
    fill = sched.n_int[0] * sched.n_sub[0]
    react = sum(sched.n_int[k] * sched.n_sub[k] for k in (1, 2, 3, 4, 7))
    return fill * (4 * F_FILL + OVH_FILL) + react * (4 * F_REACT + OVH_REACT) + F_EPILOGUE, fill + react


class ClockSampler(object):
    """nvidia-smi sampler running DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def measure_fp64_peak(core, torch, device):
    """Measured FP64 FMA-pipe peak (TFLOP/s): DFMA probe, 8 independent chains/thread, best of 5 (burst) and
    the mean over a ~1 s back-to-back loop (sustained, under the power cap)."""
    blocks, threads, iters = 148 * 16, 256, 20000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        core.fp64_probe(blocks, threads, iters, device)
    torch.cuda.synchronize()
    best, flops = 0.0, 0.0
    for _ in range(5):
        e0.record()
        _, flops = core.fp64_probe(blocks, threads, iters, device)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    reps = max(3, int(1.0 / (flops / (best * 1e12))))
    e0.record()
    for _ in range(reps):
        core.fp64_probe(blocks, threads, iters, device)
    e1.record()
    torch.cuda.synchronize()
    sustained = reps * flops / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return best, sustained


def cpu_baseline_leg(target_s):
    """cpu_baseline: the UNMODIFIED reference (baseline/_ref, installed by oracle/install_ref.py) on every host core,
    one process per core, on a bounded sample of the same workload (independent seeded SBR-v2 episodes: reset influent
    draw + one whole-cycle step) sized to ~target_s seconds; the oracle port's number is reported beside it.  Falls
    back to the port (kind "port") only when no reference install travelled with the snapshot."""
    from oracle import cpu_baseline
    cores = cpu_baseline.usable_cores()
    have_ref = cpu_baseline.reference_root() is not None
    runner = cpu_baseline.run_reference if have_ref else cpu_baseline.run
    probe = runner(steps_per_proc=2, procs=cores, warmup=1)
    per_proc_rate = probe["steps"] / cores / probe["wall_s"]
    spp = max(2, min(400, int(round(target_s * per_proc_rate))))
    r = runner(steps_per_proc=spp, procs=cores, warmup=0)
    out = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference" if have_ref else "port",
           "per_core": r["per_core"],
           "sample": "%d SBR-v2 cycle-steps (%d procs x %d: np.random.seed, SbrEnv2.reset() influent draw + "
                     "step(action) each; %s), %.1f s wall"
                     % (r["steps"], r["cores"], spp,
                        "unmodified reference from %s, scipy LSODA" % r.get("root") if have_ref
                        else "oracle port of the reference's scipy LSODA path", r["wall_s"])}
    if have_ref:
        pr = cpu_baseline.run(steps_per_proc=max(2, spp // 2), procs=cores, warmup=1)
        out["port"] = {"value": pr["value"], "per_core": pr["per_core"],
                       "note": "oracle/sbr_oracle.py restatement (no trajectory appends, no prints)"}
    return out


# interval-per-step path: algorithmic flops and bytes per env per env.step
F_EC = 94
# read: st rows 0..33 (x, controller scalars, 10-entry circular KLa history, return, steps), action, done flag;
# write: x + controller scalars + return + steps (24 rows), one KLa slot, reward, status, counters, obs_DO, obs_EC, state
OS_BYTES_STEP = 34 * 8 + 2 * 8 + 1 + 24 * 8 + 8 + 8 + 4 + 8 + (9 + 9 + 15) * 8
OS_BYTES_STEP_LEAN = OS_BYTES_STEP - 15 * 8          # emit=("obs_do", "obs_ec"): the 15-dim `state` is not written


def status_bits(torch, status):
    """How many envs carry each per-env status bit (include/sbr_b200.h SBR_ST_*)."""
    names = (("nonfinite", 1), ("waste_unassigned", 2), ("steplimit", 4), ("layers", 8))
    return {nm: int(((status & bit) != 0).sum()) for nm, bit in names}


def interval_path_leg(torch, device, args, peak_burst, peak_sustained):
    """SBROS-v1: whole episodes (reset = fill solve, 463 env.steps, the last one with settle + draw + idle) for
    --interval-envs envs, one launch per env.step, state resident in HBM.  Reports interval-steps/s and, for the
    plain react/dose steps (the dominant launch), FP64 and HBM fractions from the kernel's own RHS counters."""
    from gym_sbr2_b200 import _abi
    from gym_sbr2_b200.vec_env import SbrOsVecEnv
    n = args.interval_envs
    out = {}

    def make_actions(gen, wide):
        # per-step set-points; `moderate` is the distribution the CPU sample of this leg uses (oracle/cpu_baseline.py:
        # DO set-point U(2,3), NO3 set-point U(4,6)); `wide` (DO U(1,7), NO3 U(2,12)) doses so much carbon that 30 % of
        # the envs run into negative ammonia, where the adaptive stepper needs 5-20x more steps (a stress case)
        lo_do, w_do, lo_no, w_no = (1.0, 6.0, 2.0, 10.0) if wide else (2.0, 1.0, 4.0, 2.0)
        return [torch.stack([lo_do + w_do * torch.rand(n, dtype=torch.float64, device=device, generator=gen),
                             lo_no + w_no * torch.rand(n, dtype=torch.float64, device=device, generator=gen)], dim=1)
                for _ in range(8)]

    def episode(env, infl, acts, per_step_events):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(464)] if per_step_events else None
        rhs_mid = 0.0
        e0.record()
        env.reset(influent=infl)
        for k in range(463):
            if ev:
                ev[k].record()
            env.step_async(acts[k % 8])
            if k == 300 and ev:
                rhs_mid = env.buf.counters[0].to(torch.float64).mean()
        if ev:
            ev[463].record()
        e1.record()
        torch.cuda.synchronize()
        assert bool(env.buf.done.all()), "episode did not end after 463 steps"
        per = [ev[k].elapsed_time(ev[k + 1]) for k in range(463)] if ev else None
        return e0.elapsed_time(e1), per, float(rhs_mid)

    for mode, kw in (("dp45", dict(rtol=1e-8, atol=1e-10)), ("rk4", dict(rk4_sub_interval=20))):
        env = SbrOsVecEnv(n, device=device, seed=77, mode=mode, **kw)
        gen = torch.Generator(device=device).manual_seed(5)
        acts = make_actions(gen, wide=False)
        infl = env._draw_influent()
        env.reset(influent=infl)
        for k in range(5):
            env.step_async(acts[k % 8])
        torch.cuda.synchronize()
        ms_episode, per, rhs = episode(env, infl, acts, True)
        bad = int((env.buf.status != 0).sum())
        ret = float(env.buf.st[_abi.OS_RETURN].mean())
        plain = sorted(per[60:270] + per[280:455])
        ms_plain = plain[len(plain) // 2]
        steps = rhs / 4.0 if mode == "rk4" else (rhs - 1) / 6.0
        ovh = 196 if mode == "rk4" else 2 * 14 * (1 + 2 + 3 + 4 + 5 + 5 + 6) + 100
        flops = rhs * F_EC + steps * ovh + 150
        tf = n * flops / (ms_plain * 1e-3) / 1e12
        ms_wide, _, _ = episode(env, infl, make_actions(gen, wide=True), False)
        wide_bits = status_bits(torch, env.buf.status)
        wide_bad = int((env.buf.status != 0).sum())
        # the same plain step without the `state` output (a policy that reads obs_DO / obs_EC only), and K = 8 steps
        # per launch (frame-skip: the state stays in registers between the 8 intervals)
        env.emit = ("obs_do", "obs_ec")
        ms_lean = sorted(episode(env, infl, acts, True)[1][60:270])[105]
        env.emit = ("obs_do", "obs_ec", "state")
        env.reset(influent=infl)
        a8 = torch.stack([a.t().contiguous() for a in acts])                 # [8, 2, n]
        r8 = torch.empty((8, n), dtype=torch.float64, device=device)
        for _ in range(8):
            env.step_k(a8, r8)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            env.step_k(a8, r8)
        e1.record()
        torch.cuda.synchronize()
        ms_k8 = e0.elapsed_time(e1) / 160
        out[mode] = {"envs": n, "ms_per_episode": ms_episode, "interval_steps_per_sec": n * 463 / (ms_episode * 1e-3),
                     "actions": "per-step DO set-point U(2,3), NO3 set-point U(4,6): the distribution of the CPU sample",
                     "ms_per_plain_step": ms_plain, "plain_interval_steps_per_sec": n / (ms_plain * 1e-3),
                     "ms_terminal_step": per[462], "rhs_per_env_step": rhs,
                     "roofline": {"bound": "fp64+hbm", "fp64_tflops": tf, "fp64_frac": tf / peak_burst,
                                  "hbm_gbs": n * OS_BYTES_STEP / (ms_plain * 1e-3) / 1e9,
                                  "hbm_frac": n * OS_BYTES_STEP / (ms_plain * 1e-3) / 1e9 / _hbm_peak(),
                                  "flops_per_env_step": flops, "bytes_per_env_step": OS_BYTES_STEP},
                     "config": dict(kw, integrator=mode), "bad_status": bad,
                     "gpu_launches": 464, "mean_episode_return": ret,
                     "stress_wide_actions": {"actions": "DO set-point U(1,7), NO3 set-point U(2,12)",
                                             "ms_per_episode": ms_wide,
                                             "interval_steps_per_sec": n * 463 / (ms_wide * 1e-3),
                                             "bad_status": wide_bad, "status_bits": wide_bits},
                     "lean_outputs": {"emit": "obs_do, obs_ec (no 15-dim state)", "ms_per_plain_step": ms_lean,
                                      "bytes_per_env_step": OS_BYTES_STEP_LEAN,
                                      "hbm_frac": n * OS_BYTES_STEP_LEAN / (ms_lean * 1e-3) / 1e9 / _hbm_peak()},
                     "k8_launch": {"what": "sbr_os_step_k, K = 8 env.steps per launch (steps 64..231 of the episode)",
                                   "ms_per_env_step": ms_k8, "interval_steps_per_sec": n / (ms_k8 * 1e-3)}}
        del env
    return out


def cycle_dp45_leg(torch, device, core, env, n):
    """The same SBR-v2 batch through the adaptive Dormand-Prince mode (per-env step control): in the caller's env
    order (random set-points side by side in a warp: the divergence case) and with the divergence-aware ordering
    the vector env applies in this mode: argsort of the first set-point, one gather launch (inputs), the cycle
    kernel on unit-stride sorted buffers, one scatter launch (outputs) -- all four inside the timed region."""
    from gym_sbr2_b200 import _abi
    from gym_sbr2_b200.vec_env import SbrV2VecEnv
    res = {}
    for rtol, atol in ((1e-7, 1e-9), (1e-6, 1e-8)):
        tol = _abi.make_tol(rtol, atol)
        env_o = SbrV2VecEnv(n, device=device, seed=env.seed, mode="dp45", rtol=rtol, atol=atol, order="action")
        env_o.reset(influent=env.influent)
        for ordered in (False, True):
            def launch():
                if ordered:
                    return env_o.step_soa(env._action)
                return core.cycle_v2(env.x0, env._loading, env._action, env.params, env.sched, out=env._out,
                                     mode=_abi.MODE_DP45, tol=tol)
            launch()
            torch.cuda.synchronize()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            o = launch()
            eb.record()
            torch.cuda.synchronize()
            ms = ea.elapsed_time(eb)
            cnt = o.counters.to(torch.float64)
            rhs = float(cnt[0].mean())
            steps = (rhs - 6) / 6.0
            flops = rhs * F_REACT + steps * DP45_STEP_OVH + F_EPILOGUE
            key = "rtol%g_%s" % (rtol, "ordered" if ordered else "env_order")
            res[key] = {
                "rtol": rtol, "atol": atol, "ms": ms, "cycle_steps_per_sec": n / (ms * 1e-3),
                "rhs_per_env_mean": rhs, "rhs_per_env_max": float(cnt[0].max()),
                "warp_max_rhs_mean": float(cnt[0].view(-1, 32).max(dim=1).values.mean()) if not ordered else None,
                "rejected_per_env_mean": float(cnt[1].mean()), "bad_status": int((o.status != 0).sum()),
                "fp64_tflops": n * flops / (ms * 1e-3) / 1e12}
            if ordered:
                # the cycle kernel alone on the already sorted buffers (what the roofline fraction is quoted on)
                z = env_o._sorted
                ea.record()
                core.cycle_v2(z["x0"], z["loading"], z["action"], env_o.params, env_o.sched, out=z["out"],
                              mode=_abi.MODE_DP45, tol=tol)
                eb.record()
                torch.cuda.synchronize()
                res[key]["kernel_ms"] = ea.elapsed_time(eb)
                res[key]["kernel_fp64_tflops"] = n * flops / (res[key]["kernel_ms"] * 1e-3) / 1e12
                res[key]["launches"] = "argsort + sbr_permute_rows (gather) + sbr_cycle_v2 + sbr_permute_rows (scatter)"
            else:
                # step-count histogram (BASELINE config 3: adaptive-step divergence): quantiles of the per-env RHS
                # count, and what a warp pays for it -- the slowest of its 32 envs in env order vs in set-point order
                q = torch.tensor([0.01, 0.1, 0.5, 0.9, 0.99], dtype=torch.float64, device=device)
                sample = cnt[0][:1 << 18]
                res[key]["rhs_per_env_quantiles_p1_p10_p50_p90_p99"] = [float(v) for v in torch.quantile(sample, q)]
                by_sp = cnt[0][torch.argsort(env._action[0])]
                res[key]["warp_max_rhs_mean_if_ordered_by_first_setpoint"] = float(by_sp.view(-1, 32).max(dim=1).values.mean())
        del env_o
    return res


def v4_path_leg(torch, device, args):
    """SBR-v4 (fill phase stepped inside step(), 1-D delta-set-point action): one whole episode (493 launches)."""
    from gym_sbr2_b200 import _abi
    from gym_sbr2_b200.vec_env import SbrV4VecEnv
    n = args.interval_envs
    env = SbrV4VecEnv(n, device=device, seed=99, mode="dp45")
    gen = torch.Generator(device=device).manual_seed(6)
    acts = [0.2 * torch.randn(n, dtype=torch.float64, device=device, generator=gen) + 0.02 for _ in range(8)]
    env.reset()
    for k in range(10):                      # through the first re-sort of the state slots (step 8): its one-time
        env.step_async(acts[k % 8])          # set-up (argsort workspace, the second state buffer) is not episode time
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(495)]
    ev[0].record()
    env.reset()
    for k in range(493):
        ev[k + 1].record()
        env.step_async(acts[k % 8])
    ev[494].record()
    torch.cuda.synchronize()
    per = [ev[k + 1].elapsed_time(ev[k + 2]) for k in range(493)]
    ms_episode = ev[0].elapsed_time(ev[494])
    fill, react = sorted(per[3:26]), sorted(per[40:490])
    # the fused rollout: 8 steps per launch behind a 14 -> 32 -> 1 policy head evaluated in-kernel, every buffer in slot
    # order, slots fully re-sorted by RHS count every second launch (sbr_v4_rollout_k); the step-by-step loop with the same
    # policy [sbr_policy_mlp, sbr_v4_step] beside it
    from gym_sbr2_b200 import rollout
    policy = rollout.TinyPolicy(device, n_in=14, lo=(-0.1,), span=(0.4,), seed=3)
    fused = {}
    try:
        rollout.collect_episode_v4_fused(env, policy, K=8)                       # warm-up
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        env.epoch.zero_()
        e0.record()
        ep_f = rollout.collect_episode_v4_fused(env, policy, K=8)
        e1.record()
        env.epoch.zero_()
        ep_s = rollout.collect_episode_v4(env, policy)
        e2.record()
        torch.cuda.synchronize()
        dev_ret = float((ep_f["returns"] - ep_s["returns"]).abs().max() / ep_s["returns"].abs().max())
        fused = {"ms_episode_fused_k8": e0.elapsed_time(e1), "ms_episode_stepwise_same_policy": e1.elapsed_time(e2),
                 "interval_steps_per_sec_fused": n * 493 / (e0.elapsed_time(e1) * 1e-3),
                 "fused_vs_stepwise_max_rel_dev_of_returns": dev_ret, "all_done": bool(ep_f["all_done"]),
                 "mean_episode_return": float(ep_f["returns"].mean())}
    except Exception as exc:                                     # noqa: BLE001
        fused = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    return {"envs": n, "ms_per_episode": ms_episode, "interval_steps_per_sec": n * 493 / (ms_episode * 1e-3),
            "fused_rollout": fused,
            "ms_per_fill_step": fill[len(fill) // 2], "ms_per_react_step": react[len(react) // 2],
            "ms_terminal_step": per[492], "all_done": bool(env.buf.done.all()),
            "bad_status": int((env.buf.status != 0).sum()), "config": {"integrator": "dp45", "rtol": 1e-8, "atol": 1e-10},
            "parity_note": "against the default-tolerance reference the So component is bounded at 4 tolerance units, not 1 "
                           "(tests/test_twin_parity_v4.py V4_SO_SLACK): this env's PID has derivative action 300 x dSo per "
                           "interval, which amplifies LSODA's own noise -- the default-tolerance reference is itself up "
                           "to 1.9 units from LSODA at 1e-12 in So, this kernel < 0.05; against the 1e-12 oracle the plain "
                           "tolerance is asserted.  The oracle is the unmodified source with numpy < 1.18 linspace "
                           "semantics restored (oracle/make_golden_v4.py).",
            "gpu_launches": 494, "mean_episode_return": float(env.buf.st[_abi.V4_RETURN].mean())}


def cnt_family_leg(torch, device, args):
    """SBRCnt-v0/1/2, SBRCntMA-v1, SBROS-v2: one whole episode of each at --interval-envs envs (reset + every env.step),
    per-env DO set-points ramped to U(1, 3) g/m3 in the first aerobic steps, carbon set-point held at 0."""
    from gym_sbr2_b200 import _abi
    from gym_sbr2_b200.cnt import SbrCntVecEnv
    n = args.interval_envs
    out = {}
    for kind in ("cnt0", "cnt1", "cnt2", "ma1", "os2"):
        env = SbrCntVecEnv(kind, n, device=device, seed=77)
        gen = torch.Generator(device=device).manual_seed(8)
        target = 1.0 + 2.0 * torch.rand(n, dtype=torch.float64, device=device, generator=gen)
        zero = torch.zeros(n, dtype=torch.float64, device=device)
        steps = env.max_episode_steps
        up = range(60, 64) if kind == "ma1" else range(1, 5)

        def action(k):
            if kind == "os2":
                return torch.stack([target, zero], dim=1)
            if k == 0 and kind in ("cnt2", "ma1"):
                return zero - 2.0                       # carbon set-point 2 -> 0 before its controller can run away
            return target / 4 if k in up else zero

        acts = [action(k) for k in range(min(steps, 70))]
        env.reset()
        for k in range(3):
            env.step_async(acts[k])
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 2)]
        ev[0].record()
        env.reset()
        for k in range(steps):
            ev[k + 1].record()
            env.step_async(acts[k] if k < len(acts) else acts[-1])
        ev[steps + 1].record()
        torch.cuda.synchronize()
        per = sorted(ev[k + 1].elapsed_time(ev[k + 2]) for k in range(steps // 2, steps - 1))
        ms = ev[0].elapsed_time(ev[steps + 1])
        b = env.buf
        stats = {"all_done": bool(b.done.all()), "bad_status": int((b.status != 0).sum()),
                 "max_volume_m3": float(b.st[0].max())}
        fused = {}
        try:        # the same env behind a stand-in policy head: step by step [sbr_policy_mlp, sbr_cnt_step] and fused (K = 8)
            from gym_sbr2_b200 import rollout
            from gym_sbr2_b200.cnt import POLICY_INPUTS
            if kind == "os2":
                pol = rollout.TinyPolicy(device, n_in=18, lo=(0.5, 0.0), span=(3.0, 0.0), seed=4)
            else:
                sp = 0.004 if kind == "cnt0" else 0.06
                pol = rollout.TinyPolicy(device, n_in=POLICY_INPUTS[kind], lo=(-sp / 4,), span=(sp,), seed=4)
            rollout.collect_episode_cnt_fused(env, pol, K=8)
            f0, f1, f2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            env.epoch.zero_()
            f0.record()
            epf = rollout.collect_episode_cnt_fused(env, pol, K=8)
            f1.record()
            env.epoch.zero_()
            eps = rollout.collect_episode_cnt(env, pol)
            f2.record()
            torch.cuda.synchronize()
            fused = {"ms_per_episode_fused_k8": f0.elapsed_time(f1), "ms_per_episode_stepwise_same_policy": f1.elapsed_time(f2),
                     "all_done": bool(epf["all_done"]) and bool(eps["all_done"]),
                     "max_rel_dev_of_returns": float((epf["returns"] - eps["returns"]).abs().max()
                                                     / eps["returns"].abs().max().clamp_min(1.0))}
        except Exception as exc:                                 # noqa: BLE001
            fused = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
        out[kind] = {"behind_a_policy": fused,
                     "id": {"cnt0": "SBRCnt-v0", "cnt1": "SBRCnt-v1", "cnt2": "SBRCnt-v2", "ma1": "SBRCntMA-v1",
                            "os2": "SBROS-v2"}[kind], "envs": n, "episode_steps": steps, "ms_per_episode": ms,
                     "env_steps_per_sec": n * steps / (ms * 1e-3), "ms_per_step_median_second_half": per[len(per) // 2],
                     "all_done": stats["all_done"], "bad_status": stats["bad_status"],
                     "max_volume_m3": stats["max_volume_m3"], "gpu_launches": steps + 1}
    out["note"] = ("cnt1 / cnt2 episodes are 228 env.steps because two steps each simulate a whole anoxic phase (46 / 171 "
                   "control intervals in one solve); reward = the reference's threshold table with its unbound names bound "
                   "as oracle/make_golden_cnt.py documents")
    return out


def small_batch_leg(torch, device, n=4096):
    """BASELINE config[1] size (4096 envs): one env.step is ~10 us of GPU work, so the Python loop and launches
    dominate; reports the per-step WALL time of policy + step eagerly and through a CUDA graph (8 steps per replay),
    and the SBR-v2 whole-cycle launch at this size (a single partial wave: latency-bound)."""
    from gym_sbr2_b200 import rollout
    from gym_sbr2_b200.vec_env import SbrOsVecEnv, SbrV2VecEnv
    env = SbrOsVecEnv(n, device=device, seed=1, mode="dp45")
    pol = rollout.TinyPolicy(device)
    env.reset()
    b = env.buf
    for _ in range(5):
        env.step_soa(pol.act_into(b.obs_do, b.obs_ec, env._action))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        env.step_soa(pol.act_into(b.obs_do, b.obs_ec, env._action))
    torch.cuda.synchronize()
    t_eager = (time.perf_counter() - t0) / 200
    env.reset()
    g8 = rollout.GraphedStepper(env, pol, 8)
    env.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(25):
        g8.replay()
    torch.cuda.synchronize()
    t_graph = (time.perf_counter() - t0) / 200
    v2 = SbrV2VecEnv(n, device=device, seed=1)
    v2.reset()
    a3 = torch.rand((n, 3), dtype=torch.float64, device=device)
    v2.step(a3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        v2.step(a3)
    torch.cuda.synchronize()
    t_cycle = (time.perf_counter() - t0) / 5
    # the same 4096 cycles in the adaptive mode, and the fused SBROS-v1 rollout (8 steps per launch) at this size
    v2d = SbrV2VecEnv(n, device=device, seed=1, mode="dp45", rtol=1e-7, atol=1e-9)
    v2d.reset()
    v2d.step(a3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        v2d.step(a3)
    torch.cuda.synchronize()
    t_cycle_dp = (time.perf_counter() - t0) / 5
    env.reset()
    rollout.collect_episode_fused(env, pol, K=8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rollout.collect_episode_fused(env, pol, K=8)
    torch.cuda.synchronize()
    t_fused = (time.perf_counter() - t0) / env.max_episode_steps
    return {"envs": n, "sbros_v1_us_per_step_eager": t_eager * 1e6, "sbros_v1_us_per_step_cuda_graph": t_graph * 1e6,
            "sbros_v1_us_per_step_fused_rollout_k8": t_fused * 1e6,
            "sbros_v1_interval_steps_per_sec_cuda_graph": n / t_graph,
            "sbros_v1_interval_steps_per_sec_fused_rollout_k8": n / t_fused,
            "sbr_v2_ms_per_cycle_launch": t_cycle * 1e3, "sbr_v2_cycle_steps_per_sec": n / t_cycle,
            "sbr_v2_dp45_ms_per_cycle_launch": t_cycle_dp * 1e3, "sbr_v2_dp45_cycle_steps_per_sec": n / t_cycle_dp}


def ilc_leg(torch, device, n=1 << 16):
    """The batch-to-batch (ILC) feed-forward KLa path of `SBR-v0`: SbrIlcVecEnv.step = sbr_ilc_update (HBM-bound: seven
    [4769][N] sample-row moves per update) + sbr_cycle_ilc (FP64-bound cycle that stores So at every output point of the
    reference's grid).  Three steps; CUDA events around the whole step and, inside the same steps, around each of its two
    launches (the env's calls are wrapped for the duration of the leg)."""
    from gym_sbr2_b200 import ilc
    env = ilc.SbrIlcVecEnv(n, device=device, seed=1, learn="feedback")
    env.reset()
    a = torch.rand((n, 3), dtype=torch.float64, device=device, generator=torch.Generator(device=device).manual_seed(5)) * 4 + 0.5
    env.step(a)
    torch.cuda.synchronize()
    S = int(env.layout.n_samples)
    marks = {"update": [], "cycle": []}

    def timed(fn, key):
        def wrapper(*args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*args, **kw)
            e1.record()
            marks[key].append((e0, e1))
            return out
        return wrapper

    orig = ilc.ilc_update, ilc.cycle_ilc
    ilc.ilc_update, ilc.cycle_ilc = timed(orig[0], "update"), timed(orig[1], "cycle")
    try:
        reps, steps = 3, []
        for _ in range(reps):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            env.step(a)
            s1.record()
            steps.append((s0, s1))
        torch.cuda.synchronize()
    finally:
        ilc.ilc_update, ilc.cycle_ilc = orig
    mean = lambda ev: sum(e0.elapsed_time(e1) for e0, e1 in ev) / len(ev)
    t_step, t_up, t_cy = mean(steps), mean(marks["update"]), mean(marks["cycle"])
    rhs = float(env._cyc.counters[0].double().mean())
    bad = int((env._cyc.status != 0).sum())
    del env
    torch.cuda.empty_cache()
    # `SBR-v1`: the same plant under the feedback PID alone (the cycle kernel without feed-forward and without memories)
    v1 = ilc.SbrV1VecEnv(n, device=device, seed=1)
    v1.reset()
    v1.step(a)
    torch.cuda.synchronize()
    v1_steps = []
    for _ in range(reps):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        v1.step(a)
        s1.record()
        v1_steps.append((s0, s1))
    torch.cuda.synchronize()
    t_v1 = mean(v1_steps)
    return {"envs": n, "samples_per_env": S, "ms_per_step": t_step, "cycle_steps_per_sec": n / t_step * 1e3,
            "update_kernel_ms": t_up, "update_kernel_gbs": 7 * S * n * 8 / t_up / 1e6, "hbm_peak_gbs": _hbm_peak(),
            "cycle_kernel_ms": t_cy, "cycle_rhs_per_env": rhs, "bad_status": bad,
            "memory_gb": 6 * S * n * 8 / 1e9, "sbr_v1_ms_per_step": t_v1, "sbr_v1_cycle_steps_per_sec": n / t_v1 * 1e3,
            "note": "integrator: Dormand-Prince per PID interval (rtol 1e-9), So memory from its continuous extension; "
                    "update / cycle times are measured inside the same three steps as ms_per_step; reward by construction, "
                    "not pinned"}


def rollout_leg(torch, tdist, device, rank, world, args):
    """BASELINE config 5: --rollout-envs SBROS-v1 envs in total, sharded over the ranks by contiguous index blocks,
    one full episode (reset + 463 env.steps) driven by a small torch policy on the observation tensors, then an
    NCCL all_gather of the per-env episode returns.  Strong scaling: total work is fixed as N grows."""
    from gym_sbr2_b200 import dist, rollout
    from gym_sbr2_b200.vec_env import SbrOsVecEnv
    total = args.rollout_envs
    lo, hi = dist.shard_range(total, rank, world)
    ok, err = 1.0, ""
    try:                                                         # local set-up: no collectives in here
        env = SbrOsVecEnv(hi - lo, device=device, seed=4242, mode="dp45", env_offset=lo,    # draws keyed by GLOBAL env index
                          emit=("obs_do", "obs_ec"))                                        # all the policy reads
        policy = rollout.TinyPolicy(device)
        warm = rollout.collect_episode(env, policy, max_steps=3)     # warm-up (allocations, policy kernels)
        # the inner loop [policy -> action -> sbr_os_step] captured in CUDA graphs (8 steps and 1 step per replay):
        # at 2^20 / 8 envs per rank the eager loop is bound by Python + launches, not by the GPU
        big = rollout.GraphedStepper(env, policy, 8)
        small = rollout.GraphedStepper(env, policy, 1)
        torch.cuda.synchronize()
    except Exception as exc:                                     # noqa: BLE001
        ok, err = 0.0, "%s: %s" % (type(exc).__name__, str(exc)[:300])
    flag = torch.tensor([ok], dtype=torch.float64, device=device)
    if world > 1:
        tdist.all_reduce(flag, op=tdist.ReduceOp.MIN)            # every rank agrees before any data collective
    if float(flag.item()) < 1.0:
        return {"error": err or "set-up failed on another rank"}
    dist.gather_rewards(warm["returns"], total)                  # the NCCL channel set-up for this size
    torch.cuda.synchronize()

    def timed_episode(collect):
        if world > 1:
            tdist.barrier()
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        env.epoch.zero_()                                        # every timed episode replays the SAME influent draws
        e0.record()
        ep = collect()
        e1.record()
        if world > 1:
            tdist.barrier()                                      # absorbs the skew between ranks, so that ...
        e2.record()
        allr = dist.gather_rewards(ep["returns"], total)         # ... this times the only collective: all_gather over NVLink
        e3.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e3), e2.elapsed_time(e3), e0.elapsed_time(e1)], dtype=torch.float64,
                         device=device)
        if world > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return ep, allr, float(t[0]), float(t[1]), float(t[2])

    ep_e, _, ms_eager, _, _ = timed_episode(lambda: rollout.collect_episode(env, policy))
    ep_g, _, ms_graph, _, _ = timed_episode(lambda: rollout.collect_episode_graphed(env, big, small))
    # the fused rollout: 8 env.steps per launch with the policy head evaluated in-kernel between them (sbr_os_rollout_k);
    # same arithmetic as the step-by-step loops above, bit for bit (tested), so the returns below are theirs too
    rollout.collect_episode_fused(env, policy, K=8)              # warm-up
    ep, allr, ms_total, ms_gather, ms_local = timed_episode(lambda: rollout.collect_episode_fused(env, policy, K=8))
    _, stats = rollout.gather_episode_returns(ep["returns"], total) if world == 1 else (None, rollout.return_stats(allr))
    same = bool(torch.equal(ep["returns"], ep_g["returns"]))
    return {"total_envs": total, "envs_per_rank": hi - lo, "episode_steps": ep["steps"], "ms_episode": ms_total,
            "interval_steps_per_sec": total * ep["steps"] / (ms_total * 1e-3), "ms_reward_gather": ms_gather,
            "ms_episode_slowest_rank_before_gather": ms_local, "ms_episode_eager_loop": ms_eager,
            "ms_episode_cuda_graphs": ms_graph,
            "stepper": "fused rollout kernel sbr_os_rollout_k (8 env.steps per launch, policy head in-kernel); "
                       "ms_episode_cuda_graphs: [sbr_policy_mlp, sbr_os_step] per step captured in CUDA graphs; "
                       "ms_episode_eager_loop: the same two launches per step from Python",
            "fused_equals_stepwise_bitwise": same,
            "gathered_returns": int(allr.numel()),
            "all_done": bool(ep["all_done"]) and bool(ep_e["all_done"]) and bool(ep_g["all_done"]),
            "return_stats": stats,
            "scaling": "strong", "collective": "all_gather of per-env returns (%d B per rank)" % ((hi - lo) * 8)}


def headline_strong_leg(torch, tdist, device, rank, world, args, core, peak):
    """What north_star literally asks: --strong-envs (2^20) SBR-v2 envs IN TOTAL, sharded over the ranks by contiguous
    global index blocks (influent keyed by global index: the same 2^20 envs at every N), one whole cycle each; the
    cycle kernel's CUDA-event time (max over ranks) and its FP64 roofline fraction at this per-GPU batch."""
    from gym_sbr2_b200 import dist
    from gym_sbr2_b200.vec_env import SbrV2VecEnv
    total = args.strong_envs
    lo, hi = dist.shard_range(total, rank, world)
    n = hi - lo
    env = SbrV2VecEnv(n, device=device, seed=1234, mode=args.mode, rtol=args.rtol, atol=args.atol, env_offset=lo)
    env.reset()
    # the action of env i is a function of its GLOBAL index too
    gen = torch.Generator(device=device).manual_seed(99)
    action = torch.rand((total, 3), dtype=torch.float64, device=device, generator=gen)[lo:hi].contiguous()
    for _ in range(2):
        o = env.step_async(action)
    torch.cuda.synchronize()
    if world > 1:
        tdist.barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    ea.record()
    for _ in range(reps):
        o = env.step_async(action)
    eb.record()
    torch.cuda.synchronize()
    t = torch.tensor([ea.elapsed_time(eb) / reps], dtype=torch.float64, device=device)
    chk = torch.stack([o.reward.sum(), (o.status != 0).sum().to(torch.float64)])
    if world > 1:
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        tdist.all_reduce(chk, op=tdist.ReduceOp.SUM)
    ms = float(t.item())
    if args.mode == "rk4":
        flops_env, _ = cycle_flops_rk4(env.sched)
    else:
        cnt = o.counters.to(torch.float64)
        rhs = float(cnt[0].mean())
        flops_env = rhs * F_REACT + (rhs - 6) / 6.0 * DP45_STEP_OVH + F_EPILOGUE
    tf = n * flops_env / (ms * 1e-3) / 1e12
    ctas = (n + 63) // 64
    return {"total_envs": total, "envs_per_gpu": n, "ms_per_cycle_step": ms, "cycle_steps_per_sec": total / (ms * 1e-3),
            "scaling": "strong", "integrator": args.mode,
            "roofline": {"bound": "fp64", "achieved_per_gpu": tf, "peak": peak, "frac": tf / peak, "unit": "TFLOP/s"},
            "grid": "%d CTAs of 64 threads per GPU over 148 SMs" % ctas,
            "reward_sum_all_ranks": float(chk[0]), "bad_status": int(chk[1])}


def cycle_rk4_leg(torch, device, core, env, n, peak):
    """The headline batch through fixed-step RK4 on the reference's own output grid (9-10 sub-steps per PID interval:
    deterministic work, the highest roofline fraction, 2.3x the right-hand-side evaluations of the adaptive mode)."""
    from gym_sbr2_b200 import _abi, schedule
    sched = schedule.cycle_schedule()
    ms = []
    for _ in range(4):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        o = core.cycle_v2(env.x0, env._loading, env._action, env.params, sched, out=env._out, mode=_abi.MODE_RK4)
        eb.record()
        torch.cuda.synchronize()
        ms.append(ea.elapsed_time(eb))
    ms = sum(ms[1:]) / 3
    flops_env, steps_env = cycle_flops_rk4(sched)
    tf = n * flops_env / (ms * 1e-3) / 1e12
    return {"kernel_ms": ms, "cycle_steps_per_sec": n / (ms * 1e-3), "rhs_per_env": 4.0 * steps_env,
            "roofline": {"bound": "fp64", "achieved": tf, "peak": peak, "frac": tf / peak, "unit": "TFLOP/s",
                         "flops_per_env": flops_env}, "bad_status": int((o.status != 0).sum())}


def headline_accuracy_leg(torch, device, core, env, n):
    """How far the adaptive headline setting is from a converged solution, over the whole batch: x_last against RK4
    with 40 sub-steps per interval (4x finer than the reference grid), in units of the parity tolerance
    (1e-5 relative + 1e-9 x_1_state)."""
    from gym_sbr2_b200 import _abi, schedule
    o = env.step_soa(env._action)
    x_dp = o.x_last.clone()
    fine = core.cycle_v2(env.x0, env._loading, env._action, env.params, schedule.cycle_schedule(substeps=40),
                         mode=_abi.MODE_RK4)
    scale = torch.tensor([1.32, 30, 30, 1500, 150, 3000, 2000, 600, 8, 20, 20, 10, 10, 10], dtype=torch.float64,
                         device=device)[:, None]
    w = ((x_dp - fine.x_last).abs() / (1e-5 * fine.x_last.abs() + 1e-9 * scale)).max(dim=0).values
    return {"against": "RK4, 40 sub-steps per PID interval", "tolerance_units": {
        "median": float(w.median()), "p999": float(torch.quantile(w[:1 << 18], 0.999)), "max": float(w.max()),
        "frac_above_1": float((w > 1).double().mean())}}


def cycle_substeps_leg(torch, device, core, env, n, substeps):
    """The headline batch with fewer RK4 sub-steps per PID interval than the reference's output grid (the grid is
    a convention of the reference's trajectory logging, not of its integrator; accuracy table in
    gym_sbr2_b200/schedule.py).  NOT the headline: reported beside it."""
    from gym_sbr2_b200 import _abi, schedule
    sched = schedule.cycle_schedule(substeps=substeps)
    ref_x = env._out.x_last.clone()
    core.cycle_v2(env.x0, env._loading, env._action, env.params, env.sched, out=env._out, mode=_abi.MODE_RK4)
    torch.cuda.synchronize()
    ref_x = env._out.x_last.clone()
    core.cycle_v2(env.x0, env._loading, env._action, env.params, sched, out=env._out, mode=_abi.MODE_RK4)
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    o = core.cycle_v2(env.x0, env._loading, env._action, env.params, sched, out=env._out, mode=_abi.MODE_RK4)
    eb.record()
    torch.cuda.synchronize()
    ms = ea.elapsed_time(eb)
    scale = torch.tensor([1.32, 30, 30, 1500, 150, 3000, 2000, 600, 8, 20, 20, 10, 10, 10], dtype=torch.float64,
                         device=device)[:, None]
    w = ((o.x_last - ref_x).abs() / (1e-5 * ref_x.abs() + 1e-9 * scale)).max(dim=0).values
    return {"substeps": substeps, "kernel_ms": ms, "cycle_steps_per_sec": n / (ms * 1e-3),
            "vs_reference_grid_in_tolerance_units": {"median": float(w.median()),
                                                     "p999": float(torch.quantile(w[:1 << 18], 0.999)),
                                                     "frac_above_1": float((w > 1).double().mean())}}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path -- the unmodified package from
    baseline/_ref through its own SbrEnv2.reset()/step(), one process per host core -- same metric/unit/config.
    Falls back to the oracle port when no reference install is present.  Under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import cpu_baseline
    cores = cpu_baseline.usable_cores()
    have_ref = cpu_baseline.reference_root() is not None
    runner = cpu_baseline.run_reference if have_ref else cpu_baseline.run
    per_proc = max(1, args.ref_steps_per_proc)
    total, t_all = 0, 0.0
    for _ in range(args.warmup):
        runner(steps_per_proc=1, procs=cores, warmup=0)
    for _ in range(args.steps):
        r = runner(steps_per_proc=per_proc, procs=cores, warmup=1)
        total += r["steps"]; t_all += r["wall_s"]
    value = total / t_all
    kind = "reference" if have_ref else "port"
    sample = ("%d procs x %d SBR-v2 cycle-steps per bench step (np.random.seed, reset influent draw + step); %s"
              % (cores, per_proc, "unmodified reference (baseline/_ref) SbrEnv2, scipy LSODA" if have_ref
                 else "oracle port, scipy LSODA"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "envs_per_gpu": args.envs_per_gpu, "integrator": "lsoda"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "interval_steps_per_sec": value * INTERVALS_PER_CYCLE, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_name(args):
    return ("configs[4] at one GPU per rank: %d SBR-v2 envs per GPU, random DO-setpoint actions, per-env influent, "
            "one full cycle each" % args.envs_per_gpu)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--mode", default="dp45", choices=["rk4", "dp45"],
                    help="headline integrator: adaptive Dormand-Prince 5(4) (default) or RK4 on the reference's output grid")
    ap.add_argument("--rtol", type=float, default=None, help="DP45 tolerance (default 1e-7)")
    ap.add_argument("--atol", type=float, default=None, help="DP45 absolute tolerance per x_1_state unit (default 1e-9)")
    ap.add_argument("--ref-steps-per-proc", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work of the cpu_baseline sample")
    ap.add_argument("--no-interval-path", action="store_true", help="skip the SBROS-v1 (interval-per-step) leg")
    ap.add_argument("--interval-envs", type=int, default=1 << 20)
    ap.add_argument("--no-rollout", action="store_true", help="skip the config-5 rollout leg")
    ap.add_argument("--rollout-envs", type=int, default=1 << 20, help="TOTAL envs of the config-5 rollout (sharded)")
    ap.add_argument("--strong-envs", type=int, default=1 << 20,
                    help="TOTAL SBR-v2 envs of the strong-scaling headline leg (sharded over the ranks)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    args.rtol = 1e-7 if args.rtol is None else args.rtol
    args.atol = 1e-9 if args.atol is None else args.atol
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as tdist
    from gym_sbr2_b200 import _abi, core, dist
    from gym_sbr2_b200.vec_env import SbrV2VecEnv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device: there is no CPU fallback")
    rank, world, local = dist.init_from_env()
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch N>1 with torch.distributed.run" % (args.gpus, world))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    n = args.envs_per_gpu
    K, W = args.steps, args.warmup

    env = SbrV2VecEnv(n, device=device, seed=1234, mode=args.mode, rtol=args.rtol, atol=args.atol, env_offset=rank * n)
    env.reset()
    gen = torch.Generator(device=device).manual_seed(99 + rank)
    action = torch.rand((n, 3), dtype=torch.float64, device=device, generator=gen)
    stats = torch.empty((5,), dtype=torch.float64, device=device)

    def hot_step():
        o = env.step_async(action)
        core.reward_stats(o.reward, o.status, out=stats)
        if world > 1:
            return dist.gather_stats(stats, async_op=True)
        return stats.reshape(1, 5), None

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        g, work = hot_step()
        if work is not None:
            work.wait()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    works = []
    for _ in range(K):
        g, work = hot_step()
        if work is not None:
            works.append(work)
    for w_ in works:
        w_.wait()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        tdist.all_reduce(t_ms, op=tdist.ReduceOp.MAX)
    ms_total = float(t_ms.item())
    value = world * n * K / (ms_total * 1e-3)
    reward_stats = dist.combine_stats(g)

    # ---- dominant kernel alone: per-launch CUDA-event time of sbr_cycle_v2 on its launching stream ----------
    kern_ms = []
    # (adaptive mode: on the sorted, unit-stride buffers SbrV2VecEnv.step_soa hands it -- the argsort and the two
    # sbr_permute_rows launches around it are inside `value`, not inside this per-kernel time)
    kb = env._sorted if env.order == "action" and env._sorted is not None else None
    kx0, kload, kact, kout = (kb["x0"], kb["loading"], kb["action"], kb["out"]) if kb else \
        (env.x0, env._loading, env._action, env._out)
    for _ in range(K):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        o = core.cycle_v2(kx0, kload, kact, env.params, env.sched, out=kout, mode=env.mode, tol=env.tol)
        eb.record()
        torch.cuda.synchronize()
        kern_ms.append(ea.elapsed_time(eb))
    kern_avg = sum(kern_ms) / len(kern_ms)
    if args.mode == "rk4":
        flops_env, steps_env = cycle_flops_rk4(env.sched)
        rhs_mean, rej_mean = 4.0 * steps_env, 0.0
    else:
        cnt = o.counters.to(torch.float64)
        rhs_mean, rej_mean = float(cnt[0].mean()), float(cnt[1].mean())
        n_phases = 6                                   # FSAL restarts once per phase, not per interval
        dp_steps = (rhs_mean - n_phases) / 6.0
        flops_env = rhs_mean * F_REACT + dp_steps * DP45_STEP_OVH + F_EPILOGUE   # fill RHS counted as react: conservative
    achieved_tf = n * flops_env / (kern_avg * 1e-3) / 1e12
    peak_burst, peak_sustained = measure_fp64_peak(core, torch, device)
    peak = peak_sustained if kern_avg * K > 500 else peak_burst
    sm_mhz = torch.cuda.get_device_properties(device).clock_rate / 1e3 if hasattr(
        torch.cuda.get_device_properties(device), "clock_rate") else None
    try:
        sm_mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"])
    except Exception:
        pass
    n_sm = torch.cuda.get_device_properties(device).multi_processor_count
    peak_theory = n_sm * 64 * 2 * (sm_mhz or 1965.0) * 1e6 / 1e12       # SMs x 64 DFMA/clk x 2 flop x max SM clock
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        try:
            traffic = json.load(open(tj)).get("sbr_cycle_v2_dp45_bytes_per_env" if args.mode == "dp45" else
                                              "sbr_cycle_v2_bytes_per_env") * n
        except Exception:
            traffic = None
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tf / peak,
                "traffic": traffic, "kernel": "sbr_cycle_v2_kernel<%s>" % args.mode, "kernel_ms": kern_avg,
                "flops_per_env": flops_env, "rhs_per_env": rhs_mean, "rejected_per_env": rej_mean,
                "peak_source": "in-run DFMA probe (MEASURED_PEAKS.json has no FP64 entry); burst %.2f, sustained %.2f"
                               % (peak_burst, peak_sustained),
                "peak_theoretical": peak_theory, "frac_of_theoretical": achieved_tf / peak_theory,
                "peak_theoretical_source": "%d SMs x 64 DFMA/clk x 2 flop x %.0f MHz (max SM clock)"
                                           % (n_sm, sm_mhz or 1965.0),
                "hbm": {"bytes_per_env": BYTES_PER_ENV, "achieved_gbs": n * BYTES_PER_ENV / (kern_avg * 1e-3) / 1e9,
                        "peak_gbs": _hbm_peak()}}

    # ---- end to end through the public API with HOST buffers ---------------------------------------------------
    act_h = torch.rand((n, 3), dtype=torch.float64).pin_memory()
    infl_h = env.influent.cpu().pin_memory()
    obs_h = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    rew_h = torch.empty((n,), dtype=torch.float64).pin_memory()
    done_h = torch.empty((n,), dtype=torch.bool).pin_memory()
    act_d = torch.empty((n, 3), dtype=torch.float64, device=device)
    infl_d = torch.empty((14, n), dtype=torch.float64, device=device)

    def e2e_step():
        infl_d.copy_(infl_h, non_blocking=True)
        act_d.copy_(act_h, non_blocking=True)
        env.reset(influent=infl_d)
        obs, reward, done, info = env.step(act_d)
        obs_h.copy_(obs, non_blocking=True)
        rew_h.copy_(reward, non_blocking=True)
        done_h.copy_(done, non_blocking=True)
        torch.cuda.synchronize()

    # Pipelined variant of the same loop: step k+1's host->device copies run on a side stream while step k computes
    # (double-buffered device inputs, events both ways); every step still uploads its inputs and downloads its results
    # inside the timed region -- only the waiting is overlapped.
    copy_s = torch.cuda.Stream(device=device)
    infl_db = [torch.empty_like(infl_d) for _ in range(2)]
    act_db = [torch.empty_like(act_d) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def e2e_pipelined(steps):
        cur = torch.cuda.current_stream()
        for b in range(2):
            ev_free[b].record(cur)

        def upload(k):
            b = k & 1
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(ev_free[b])
                infl_db[b].copy_(infl_h, non_blocking=True)
                act_db[b].copy_(act_h, non_blocking=True)
                ev_in[b].record(copy_s)
        upload(0)
        for k in range(steps):
            if k + 1 < steps:
                upload(k + 1)
            b = k & 1
            cur.wait_event(ev_in[b])
            env.reset(influent=infl_db[b])
            obs, reward, done, info = env.step(act_db[b])
            ev_free[b].record(cur)
            obs_h.copy_(obs, non_blocking=True)
            rew_h.copy_(reward, non_blocking=True)
            done_h.copy_(done, non_blocking=True)
        torch.cuda.synchronize()

    def timed(fn):
        barrier()
        t0 = time.perf_counter()
        fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(2):
        e2e_step()
    e2e_pipelined(2)
    t_sync = timed(lambda: [e2e_step() for _ in range(K)])
    t_pipe = timed(lambda: e2e_pipelined(K))
    e2e = {"value": world * n * K / t_pipe, "unit": UNIT,
           "value_synchronous": world * n * K / t_sync,
           "h2d_bytes_per_step": world * (act_h.numel() + infl_h.numel()) * 8,
           "d2h_bytes_per_step": world * ((obs_h.numel() + rew_h.numel()) * 8 + done_h.numel()),
           "api": "SbrV2VecEnv.reset(influent) + step(action), pinned host buffers; every step uploads its inputs and "
                  "reads its results back; `value`: uploads of step k+1 overlap step k on a side stream, "
                  "`value_synchronous`: copy-in, step, copy-out, synchronize"}

    # legs beside the headline: a failure in one of them must not cost the headline line
    def leg(fn, *a):
        try:
            return fn(*a)
        except Exception as exc:          # noqa: BLE001 - reported, not swallowed
            return {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}

    def cpu_os_leg():
        from oracle import cpu_baseline as _cb
        have_ref = _cb.reference_root() is not None
        r = _cb.run_reference(steps_per_proc=6000, warmup=20, path="os") if have_ref else _cb.run_os(steps_per_proc=6000)
        return {"value": r["value"], "unit": "interval-steps/s", "cores": r["cores"],
                "kind": "reference" if have_ref else "port", "per_core": r["per_core"],
                "sample": "%d SBROS-v1 env.steps (%d procs x 6000; %s), %.1f s wall"
                          % (r["steps"], r["cores"], "unmodified reference SbrOS, scipy LSODA" if have_ref
                             else "oracle port of the reference's scipy LSODA path", r["wall_s"])}

    cpu, cpu_os = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = leg(cpu_baseline_leg, args.cpu_seconds)
        cpu_os = leg(cpu_os_leg)

    paths = {}
    if rank == 0 and world == 1 and not args.no_interval_path:
        paths["sbros_v1"] = leg(interval_path_leg, torch, device, args, peak_burst, peak_sustained)
        paths["sbros_v1"]["cpu_baseline"] = cpu_os
        # the other integrator settings on the same batch: adaptive at two tolerances in env order and ordered, RK4 on the
        # reference's own grid (round 1's headline) and with 7 sub-steps per interval
        paths["sbr_v2_dp45"] = leg(cycle_dp45_leg, torch, device, core, env, n)
        paths["sbr_v2_rk4_reference_grid"] = leg(cycle_rk4_leg, torch, device, core, env, n, peak_burst)
        paths["sbr_v2_rk4_7substeps"] = leg(cycle_substeps_leg, torch, device, core, env, n, 7)
        if args.mode == "dp45":
            paths["headline_accuracy"] = leg(headline_accuracy_leg, torch, device, core, env, n)
        paths["sbr_v4"] = leg(v4_path_leg, torch, device, args)
        paths["sbr_cnt_family"] = leg(cnt_family_leg, torch, device, args)
        paths["config1_small_batch"] = leg(small_batch_leg, torch, device)
        paths["sbr_v0_ilc"] = leg(ilc_leg, torch, device)
    if not args.no_rollout:
        paths["headline_strong_scaling"] = (leg(headline_strong_leg, torch, tdist, device, rank, world, args, core, peak)
                                            if world == 1 else
                                            headline_strong_leg(torch, tdist, device, rank, world, args, core, peak))
        # every rank takes part: envs sharded over the ranks, NCCL gather of the episode returns
        # (guarded only on one GPU: with several ranks a swallowed exception on one of them would leave the others
        # waiting in a collective -- there a failure must bring the job down)
        paths["config5_rollout"] = (leg(rollout_leg, torch, tdist, device, rank, world, args) if world == 1
                                    else rollout_leg(torch, tdist, device, rank, world, args))

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args), "envs_per_gpu": n, "integrator": args.mode,
                           "rtol": args.rtol if args.mode == "dp45" else None,
                           "l2": "inputs %d MB per step > 126 MB L2" % (n * 31 * 8 >> 20),
                           "parallelism": "env-sharded x%d, no step-path collective" % world},
                "interval_steps_per_sec": value * INTERVALS_PER_CYCLE,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": (5 if kb else 3) * K,
                "gpu_launch_names": ["sbr_cycle_v2_kernel", "sbr_reward_stats_init_kernel", "sbr_reward_stats_kernel"]
                + (["sbr_permute_rows_kernel (gather)", "sbr_permute_rows_kernel (scatter)"] if kb else []),
                "clocks": clocks, "reward_stats": reward_stats, "paths": paths}
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.destroy_process_group()
    return 0


def _hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


if __name__ == "__main__":
    sys.exit(main())
