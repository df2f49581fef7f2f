"""Host-side time schedule of one SBR cycle.

The reference derives every loop count from `int(float_expr)` truncation:
  * phase boundaries: phase 1 starts at 0, every later phase at previous end + t_delta, length
    t_cycle * t_ratio[k]                                   (SBR_model_FB.py:17-25, 60-264)
  * PID intervals per phase: len(linspace(t0, t1, int((t1-t0)/(10*t_delta)))) - 1
                                                            (sub_phases_FB.py:183-184)
  * output points per interval: int((te-ts)/t_delta)        (sub_phases_FB.py:231)
so the schedule is computed here with the same numpy float64 expressions and shipped to the kernel as a small
table (SbrSchedule) instead of being re-derived on the device, where FMA contraction could change a count.
"""
import numpy as np

from . import _abi

T_CYCLE = 12 / 24
T_RATIO = (4.2 / 100, 8.3 / 100, 37.5 / 100, 31.2 / 100, 2.1 / 100, 8.3 / 100, 2.1 / 100, 6.3 / 100)
T_DELTA = 0.002 / 24


def phase_bounds(t_cycle=T_CYCLE, t_ratio=T_RATIO, t_delta=T_DELTA):
    bounds = []
    t_end = 0
    for k in range(8):
        t_start = t_end if k == 0 else t_end + t_delta
        t_end = t_start + t_cycle * t_ratio[k]
        bounds.append((t_start, t_end))
    return bounds


def pid_grid(t_start, t_end, t_delta=T_DELTA):
    """(interval boundaries, output points per interval) of one PID-controlled phase."""
    t_save2 = np.linspace(t_start, t_end, int((t_end - t_start) / (t_delta * 10)))
    pts = [int((t_save2[i + 1] - t_save2[i]) / t_delta) for i in range(len(t_save2) - 1)]
    return t_save2, pts


def cycle_schedule(t_cycle=T_CYCLE, t_ratio=T_RATIO, t_delta=T_DELTA, substeps=None):
    """SbrSchedule for the cycle-per-step path (SBR-v2).

    substeps: RK4 sub-steps per PID interval.  None (default) = the reference's own output grid (9, phase 5: 10) --
    the reference's LSODA does not step on that grid, it only interpolates onto it, so this is a convention, not a
    requirement.  Measured on 4096 random envs against RK4 with 40 sub-steps, in units of the parity tolerance
    (rtol 1e-5): reference grid median 0.004 / 99.9 % 0.056; 8 sub-steps 0.009 / 0.098; 7 sub-steps 0.015 / 0.17
    (-22 % work); 5 in the anoxic phases only 0.005 / 0.41; 4 anywhere breaks the tolerance.  The reference's own
    default-tolerance LSODA sits ~0.26 units from the converged solution."""
    s = _abi.SbrSchedule()
    bounds = phase_bounds(t_cycle, t_ratio, t_delta)
    for k in range(8):
        t0, t1 = bounds[k]
        if k in (5, 6):
            s.n_int[k], s.n_sub[k], s.interval[k] = 0, 0, 0.0
            continue
        grid, pts = pid_grid(t0, t1, t_delta)
        n_int = len(grid) - 1
        if n_int < 1:
            raise ValueError("phase %d has no PID interval" % (k + 1))
        if len(set(pts)) != 1:
            raise ValueError("phase %d: output points per interval are not uniform: %s" % (k + 1, sorted(set(pts))))
        s.n_int[k] = n_int
        s.n_sub[k] = pts[0] - 1 if substeps is None else int(substeps)   # default: gaps between the reference's output points
        s.interval[k] = (t1 - t0) / n_int
    # the settler integrates over linspace(t0, t1, int((t1-t0)/t_delta)) -> total span t1 - t0
    s.settle_time = bounds[5][1] - bounds[5][0]
    return s


def schedule_summary(s):
    return dict(n_int=list(s.n_int), n_sub=list(s.n_sub), interval=list(s.interval), settle_time=s.settle_time)


# ---------------------------------------------------------------------------------------------------------
# interval-per-step path (SBROS-v1)
# ---------------------------------------------------------------------------------------------------------
def batch_time_marks(t_cycle=T_CYCLE, t_ratio=T_RATIO, t_delta=T_DELTA * 10):
    """First and last output stamp of each phase, [(t_memoryK[0], t_memoryK[-1])] for K = 1..8, as
    module_batch_time.batch_time builds them (module_batch_time.py:3-116) when called with t_delta = 10*dt
    (gym_SBR_oneshot.py:35-36): per phase a linspace of int(len/(10*t_delta)) stamps, each gap refilled with
    int(gap/t_delta) points -- every count an int(float) truncation."""
    marks = []
    t_end = 0
    for k in range(8):
        t_start = t_end if k == 0 else t_end + t_delta
        t_end = t_start + t_cycle * t_ratio[k]
        t_save = np.linspace(t_start, t_end, int((t_end - t_start) / (t_delta * 10)))
        first, last = t_save[0], t_save[0]
        for i in range(len(t_save) - 1):
            t_range = np.linspace(t_save[i], t_save[i + 1], int((t_save[i + 1] - t_save[i]) / t_delta))
            if len(t_range) > 1:
                last = t_range[-1]
        marks.append((float(first), float(last)))
    return marks


def phase_stamps(t_cycle=T_CYCLE, t_ratio=T_RATIO, t_delta=T_DELTA):
    """The full output-stamp list t_memoryK of each phase, K = 1..8 (module_batch_time.py:3-116): the phase start, then the
    int(gap / t_delta) - 1 later points of every PID interval.  These are the sample times of the batch-to-batch
    controller's memories (gym_SBR_env0.py:48, module_batch_PID.py:32-35)."""
    out = []
    t_end = 0
    for k in range(8):
        t_start = t_end if k == 0 else t_end + t_delta
        t_end = t_start + t_cycle * t_ratio[k]
        t_save = np.linspace(t_start, t_end, int((t_end - t_start) / (t_delta * 10)))
        mem = [float(t_save[0])]
        for i in range(len(t_save) - 1):
            t_range = np.linspace(t_save[i], t_save[i + 1], int((t_save[i + 1] - t_save[i]) / t_delta))
            mem.extend(float(v) for v in t_range[1:])
        out.append(mem)
    return out


def os_schedule(t_cycle=T_CYCLE, t_ratio=T_RATIO, dt=T_DELTA, rk4_sub_interval=0, rk4_sub_fill=0, rk4_sub_idle=0):
    """SbrOsSchedule: the time constants SbrOS.step / reset key on (gym_SBR_oneshot.py:28-36, 292, 860-963, 1122,
    2264-2420, 2554-2597)."""
    t_delta = dt * 10
    m = batch_time_marks(t_cycle, t_ratio, t_delta)
    s = _abi.SbrOsSchedule()
    s.tm3_0, s.tm3_1, s.tm4_1, s.tm5_1 = m[2][0], m[2][1], m[3][1], m[4][1]
    s.dt, s.t_delta = dt, t_delta
    s.t_fill = 0 + t_ratio[0] * 0.5                      # Sim_filling(x0, t=0, t_ratio[0], ...) (:292,1585)
    s.fill_pts = int((s.t_fill - 0) / dt)
    s.settle_len = t_ratio[5] * t_cycle
    s.draw_len = t_ratio[6] * t_cycle
    s.t_cycle = t_cycle
    s.rk4_sub_interval, s.rk4_sub_fill, s.rk4_sub_idle = int(rk4_sub_interval), int(rk4_sub_fill), int(rk4_sub_idle)
    return s


def os_fill_flow(qin, t_cycle=T_CYCLE, t_ratio=T_RATIO, dt=T_DELTA):
    """influent_mixed[0] := Qin / t_memory1[-1] (gym_SBR_oneshot.py:287)."""
    return qin / batch_time_marks(t_cycle, t_ratio, dt * 10)[0][1]
