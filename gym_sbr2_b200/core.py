"""Torch-facing wrappers of the C ABI (include/sbr_b200.h).  PyTorch is plumbing here: it owns device memory
and streams; all arithmetic happens in the sm_100a kernels of libsbr_b200.so.  No CPU fallback -- tensors that
are not CUDA float64 SoA are rejected."""
import ctypes as C

import torch

from . import _abi


def _dev_ptr(t, rows, n, dtype=torch.float64, name="tensor"):
    if t is None:
        return None, n
    if not t.is_cuda:
        raise _abi.SbrLibraryError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if rows == 1:
        if t.dim() != 1 or t.shape[0] != n or t.stride(0) != 1:
            raise ValueError("%s must be a contiguous [%d] vector" % (name, n))
        return C.c_void_p(t.data_ptr()), n
    if t.dim() != 2 or t.shape[0] != rows or t.shape[1] != n or (n > 1 and t.stride(1) != 1):
        raise ValueError("%s must be SoA [%d, %d] with unit stride along envs, got %s" % (name, rows, n, tuple(t.shape)))
    return C.c_void_p(t.data_ptr()), t.stride(0)


def soa1(t):
    """Normalise a [rows, 1] tensor to row stride 1 (a fresh buffer if needed)."""
    if t is not None and t.dim() == 2 and t.shape[1] == 1 and t.stride(0) != 1:
        out = torch.empty((t.shape[0], 1), dtype=t.dtype, device=t.device)
        out.copy_(t)
        return out
    return t


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _same_ld(lds, what):
    lds = {ld for ld in lds if ld is not None}
    if len(lds) != 1:
        raise ValueError("%s: all SoA tensors must share one row stride, got %s" % (what, sorted(lds)))
    return lds.pop()


class CycleV2Out(object):
    """Output buffers of sbr_cycle_v2; allocated once and reused by the env."""

    def __init__(self, n, device):
        f = dict(dtype=torch.float64, device=device)
        self.x_last = torch.empty((_abi.NX, n), **f)
        self.obs = torch.empty((3, n), **f)
        self.reward = torch.empty((n,), **f)
        self.aux = torch.empty((_abi.AUX_ROWS, n), **f)
        self.status = torch.empty((n,), dtype=torch.int32, device=device)
        self.counters = torch.empty((2, n), dtype=torch.int32, device=device)


def cycle_v2(x0, influent, action, params, sched, out=None, mode=_abi.MODE_RK4, tol=None, stream=None, perm=None):
    """One whole cycle for a batch (SbrEnv2.step, gym_SBR_env2.py:131-171).  x0, influent [14,n]; action [3,n];
    perm: optional int64 [n] permutation -- thread i works on env perm[i] (divergence-aware ordering for DP45)."""
    lib = _abi.load()
    n = x0.shape[1]
    if out is None:
        out = CycleV2Out(n, x0.device)
    px0, l0 = _dev_ptr(x0, _abi.NX, n, name="x0")
    pin, l1 = _dev_ptr(influent, _abi.NX, n, name="influent")
    pac, l2 = _dev_ptr(action, 3, n, name="action")
    pxl, l3 = _dev_ptr(out.x_last, _abi.NX, n, name="x_last")
    pob, l4 = _dev_ptr(out.obs, 3, n, name="obs")
    prw, _ = _dev_ptr(out.reward, 1, n, name="reward")
    pax, l5 = _dev_ptr(out.aux, _abi.AUX_ROWS, n, name="aux")
    pst, _ = _dev_ptr(out.status, 1, n, dtype=torch.int32, name="status")
    pct, l6 = _dev_ptr(out.counters, 2, n, dtype=torch.int32, name="counters")
    ld = _same_ld([l0, l1, l2, l3, l4, l5, l6], "cycle_v2")
    tol = tol or _abi.make_tol()
    ppm, _ = _dev_ptr(perm, 1, n, dtype=torch.int64, name="perm")
    with torch.cuda.device(x0.device):
        rc = lib.sbr_cycle_v2(n, ld, px0, pin, pac, C.byref(params), C.byref(sched), pxl, pob, prw, pax, pst, pct,
                              int(mode), C.byref(tol), ppm, _stream_ptr(stream))
    _abi.check(rc, "sbr_cycle_v2")
    return out


def cycle_v2_traj(x0, influent, action, params, sched, t_start, traj=None, out=None, mode=_abi.MODE_RK4, tol=None, stream=None):
    """sbr_cycle_v2_traj: one whole cycle with a record [t, x[14], KLa] at the end of every PID interval (+ the post-draw
    state) -- SBR_model_FB.run's `t`, `x` and KLa arrays sampled at the interval ends.  traj [R,16,n]; returns (out, traj)."""
    lib = _abi.load()
    n = x0.shape[1]
    if out is None:
        out = CycleV2Out(n, x0.device)
    R = int(lib.sbr_cycle_v2_traj_records(C.byref(sched)))
    if traj is None:
        traj = torch.full((R, _abi.TRAJ2_ROWS, n), float("nan"), dtype=torch.float64, device=x0.device)
    if tuple(traj.shape) != (R, _abi.TRAJ2_ROWS, n) or not traj.is_contiguous() or traj.dtype != torch.float64 or not traj.is_cuda:
        raise ValueError("traj must be a contiguous CUDA float64 [%d, %d, %d] tensor" % (R, _abi.TRAJ2_ROWS, n))
    px0, l0 = _dev_ptr(x0, _abi.NX, n, name="x0")
    pin, l1 = _dev_ptr(influent, _abi.NX, n, name="influent")
    pac, l2 = _dev_ptr(action, 3, n, name="action")
    pxl, l3 = _dev_ptr(out.x_last, _abi.NX, n, name="x_last")
    pob, l4 = _dev_ptr(out.obs, 3, n, name="obs")
    prw, _ = _dev_ptr(out.reward, 1, n, name="reward")
    pax, l5 = _dev_ptr(out.aux, _abi.AUX_ROWS, n, name="aux")
    pst, _ = _dev_ptr(out.status, 1, n, dtype=torch.int32, name="status")
    pct, l6 = _dev_ptr(out.counters, 2, n, dtype=torch.int32, name="counters")
    ld = _same_ld([l0, l1, l2, l3, l4, l5, l6, n], "cycle_v2_traj")
    ts = (C.c_double * _abi.NPHASE)(*[float(v) for v in t_start])
    tol = tol or _abi.make_tol()
    with torch.cuda.device(x0.device):
        rc = lib.sbr_cycle_v2_traj(n, ld, px0, pin, pac, C.byref(params), C.byref(sched), ts, pxl, pob, prw, pax, pst, pct,
                                   C.c_void_p(traj.data_ptr()), int(mode), C.byref(tol), _stream_ptr(stream))
    _abi.check(rc, "sbr_cycle_v2_traj")
    return out, traj


def integrate_interval(x, kla, params, tail, T, n_sub, mode=_abi.MODE_RK4, tol=None, ec=None, loading=None,
                       counters=None, stream=None):
    """In-place advance of x [14,n] over one interval (replaces one odeint call, sub_phases_FB.py:252,480)."""
    lib = _abi.load()
    n = x.shape[1]
    px, l0 = _dev_ptr(x, _abi.NX, n, name="x")
    pk, _ = _dev_ptr(kla, 1, n, name="kla")
    pe, _ = _dev_ptr(ec, 1, n, name="ec")
    pl, l1 = _dev_ptr(loading, _abi.NX, n, name="loading")
    pc, l2 = _dev_ptr(counters, 2, n, dtype=torch.int32, name="counters")
    ld = _same_ld([l0, l1 if loading is not None else None, l2 if counters is not None else None],
                  "integrate_interval")
    tol = tol or _abi.make_tol()
    with torch.cuda.device(x.device):
        rc = lib.sbr_integrate_interval(n, ld, px, pk, pe, pl, C.byref(params), int(tail), float(T), int(n_sub),
                                        int(mode), C.byref(tol), pc, _stream_ptr(stream))
    _abi.check(rc, "sbr_integrate_interval")
    return x


def rhs(x, kla, params, tail, ec=None, loading=None, stream=None):
    """dx = f(x): the kinetic right-hand side (rxn.dxdt / filling.dxdt / reaction_dxdt)."""
    lib = _abi.load()
    n = x.shape[1]
    dx = torch.zeros_like(x)
    px, l0 = _dev_ptr(x, _abi.NX, n, name="x")
    pd, l1 = _dev_ptr(dx, _abi.NX, n, name="dx")
    pk, _ = _dev_ptr(kla, 1, n, name="kla")
    pe, _ = _dev_ptr(ec, 1, n, name="ec")
    pl, l2 = _dev_ptr(loading, _abi.NX, n, name="loading")
    ld = _same_ld([l0, l1, l2 if loading is not None else None], "rhs")
    with torch.cuda.device(x.device):
        rc = lib.sbr_rhs(n, ld, px, pk, pe, pl, C.byref(params), int(tail), pd, _stream_ptr(stream))
    _abi.check(rc, "sbr_rhs")
    return dx


def settle_draw(x, params, settle_time, stream=None):
    """Settle + draw stage on x [14,n] IN PLACE (sub_phases_FB.py:716-915).  Returns (sX [10,n], out [9,n] = Xf, Qw,
    EQI, eff[6], status [n])."""
    lib = _abi.load()
    n = x.shape[1]
    f = dict(dtype=torch.float64, device=x.device)
    sX, out = torch.empty((10, n), **f), torch.empty((9, n), **f)
    status = torch.empty((n,), dtype=torch.int32, device=x.device)
    px, l0 = _dev_ptr(x, _abi.NX, n, name="x")
    ps, l1 = _dev_ptr(sX, 10, n, name="sX")
    po, l2 = _dev_ptr(out, 9, n, name="out")
    pst, _ = _dev_ptr(status, 1, n, dtype=torch.int32, name="status")
    ld = _same_ld([l0, l1, l2], "settle_draw")
    with torch.cuda.device(x.device):
        rc = lib.sbr_settle_draw(n, ld, px, float(settle_time), C.byref(params), ps, po, pst, _stream_ptr(stream))
    _abi.check(rc, "sbr_settle_draw")
    return sX, out, status


class OsBuffers(object):
    """Device buffers of the interval-per-step path: persistent state + per-step outputs (allocated once)."""

    def __init__(self, n, device):
        f = dict(dtype=torch.float64, device=device)
        self.st = torch.zeros((_abi.OS_ROWS, n), **f)
        self.obs_do = torch.empty((_abi.OS_NOBS, n), **f)
        self.obs_ec = torch.empty((_abi.OS_NOBS, n), **f)
        self.state = torch.empty((_abi.OS_NSTATE, n), **f)
        self.reward = torch.zeros((n,), **f)
        self.done = torch.ones((n,), dtype=torch.uint8, device=device)     # nothing to step before reset
        self.status = torch.zeros((n,), dtype=torch.int32, device=device)
        self.counters = torch.zeros((2, n), dtype=torch.int32, device=device)


def os_reset(buf, influent, params, sched, x0=None, mask=None, mode=_abi.MODE_DP45, tol=None, stream=None):
    """Episode start for a batch (SbrOS.reset + Sim_filling, gym_SBR_oneshot.py:168-438, 1585-1654).
    influent [14,n] (row 0 = fill flow); x0 [14,n] or None (reference x0_init); mask [n] uint8 or None."""
    lib = _abi.load()
    n = buf.st.shape[1]
    pst, l0 = _dev_ptr(buf.st, _abi.OS_ROWS, n, name="st")
    pin, l1 = _dev_ptr(influent, _abi.NX, n, name="influent")
    px0, l2 = _dev_ptr(x0, _abi.NX, n, name="x0")
    pmk, _ = _dev_ptr(mask, 1, n, dtype=torch.uint8, name="mask")
    pod, l3 = _dev_ptr(buf.obs_do, _abi.OS_NOBS, n, name="obs_do")
    poe, l4 = _dev_ptr(buf.obs_ec, _abi.OS_NOBS, n, name="obs_ec")
    pdn, _ = _dev_ptr(buf.done, 1, n, dtype=torch.uint8, name="done")
    pss, _ = _dev_ptr(buf.status, 1, n, dtype=torch.int32, name="status")
    pct, l5 = _dev_ptr(buf.counters, 2, n, dtype=torch.int32, name="counters")
    ld = _same_ld([l0, l1, l2 if x0 is not None else None, l3, l4, l5], "os_reset")
    tol = tol or _abi.make_tol()
    with torch.cuda.device(buf.st.device):
        rc = lib.sbr_os_reset(n, ld, px0, pin, pmk, C.byref(params), C.byref(sched), pst, pod, poe, pdn, pss, pct,
                              int(mode), C.byref(tol), _stream_ptr(stream))
    _abi.check(rc, "sbr_os_reset")
    return buf


def os_step(buf, action, params, sched, mode=_abi.MODE_DP45, tol=None, stream=None, emit=("obs_do", "obs_ec", "state"),
            rewards=None, traj=None):
    """One env.step for a batch (SbrOS.step, gym_SBR_oneshot.py:843-1273).  action [2,n]: DO and NO3 set-points --
    or [K,2,n] with rewards [K,n]: K consecutive steps in one launch (sbr_os_step_k).  emit: which of the
    observation outputs are written (the others cost no memory traffic and keep their old contents).
    traj: optional [cap, TRAJ_ROWS, n] float64 buffer receiving one record per PID interval (sbr_os_step_traj)."""
    lib = _abi.load()
    n = buf.st.shape[1]
    pst, l0 = _dev_ptr(buf.st, _abi.OS_ROWS, n, name="st")
    ptj, cap, l7 = None, 0, None
    if traj is not None:
        if traj.dim() != 3 or traj.shape[1] != _abi.TRAJ_ROWS or traj.shape[2] != n or not traj.is_contiguous():
            raise ValueError("traj must be contiguous [cap, %d, n]" % _abi.TRAJ_ROWS)
        cap = traj.shape[0]
        ptj, l7 = _dev_ptr(traj.view(cap * _abi.TRAJ_ROWS, n), cap * _abi.TRAJ_ROWS, n, name="traj")
    if action.dim() == 3:
        K = action.shape[0]
        if rewards is None or rewards.shape != (K, n):
            raise ValueError("K-step launch: rewards must be [K, n]")
        if action.shape[1:] != (2, n) or not action.is_contiguous() or not rewards.is_contiguous():
            raise ValueError("K-step launch: action must be contiguous [K, 2, n], rewards contiguous [K, n]")
        pac, l1 = _dev_ptr(action.view(2 * K, n), 2 * K, n, name="action")
        prw, l6 = _dev_ptr(rewards, K, n, name="rewards") if K > 1 else (C.c_void_p(rewards.data_ptr()), n)
        if K == 1:
            l1 = n
    else:
        K = 1
        pac, l1 = _dev_ptr(action, 2, n, name="action")
        prw, _ = _dev_ptr(buf.reward, 1, n, name="reward")
        l6 = None
    pod, l2 = _dev_ptr(buf.obs_do if "obs_do" in emit else None, _abi.OS_NOBS, n, name="obs_do")
    poe, l3 = _dev_ptr(buf.obs_ec if "obs_ec" in emit else None, _abi.OS_NOBS, n, name="obs_ec")
    pse, l4 = _dev_ptr(buf.state if "state" in emit else None, _abi.OS_NSTATE, n, name="state")
    pdn, _ = _dev_ptr(buf.done, 1, n, dtype=torch.uint8, name="done")
    pss, _ = _dev_ptr(buf.status, 1, n, dtype=torch.int32, name="status")
    pct, l5 = _dev_ptr(buf.counters, 2, n, dtype=torch.int32, name="counters")
    ld = _same_ld([l0, l1, l2 if pod else None, l3 if poe else None, l4 if pse else None, l5, l6, l7], "os_step")
    tol = tol or _abi.make_tol()
    with torch.cuda.device(buf.st.device):
        rc = lib.sbr_os_step_traj(n, ld, K, pst, pac, C.byref(params), C.byref(sched), pod, poe, pse, prw, pdn, pss,
                                  pct, int(mode), C.byref(tol), ptj, int(cap), _stream_ptr(stream))
    _abi.check(rc, "sbr_os_step")
    return buf


def os_rollout_k(buf, action, policy, rewards, params, sched, mode=_abi.MODE_DP45, tol=None, stream=None,
                 emit=("obs_do", "obs_ec"), act_log=None, obs_log=None):
    """K = rewards.shape[0] consecutive env.steps in ONE launch with the policy head evaluated in-kernel between them
    (sbr_os_rollout_k).  action [2,n] in/out: the set-points of the first step in, those of the step after the last one
    out.  policy: an _abi.SbrPolicyMlp (18 -> hidden -> 2).  rewards [K,n] out.  act_log [K,2,n] / obs_log [K,18,n]:
    optional per-step records for the learner."""
    lib = _abi.load()
    n = buf.st.shape[1]
    K = rewards.shape[0]
    if rewards.shape != (K, n) or not rewards.is_contiguous():
        raise ValueError("rewards must be contiguous [K, n]")
    pst, l0 = _dev_ptr(buf.st, _abi.OS_ROWS, n, name="st")
    pac, l1 = _dev_ptr(action, 2, n, name="action")
    prw, l2 = _dev_ptr(rewards, K, n, name="rewards") if K > 1 else (C.c_void_p(rewards.data_ptr()), None)
    pod, l3 = _dev_ptr(buf.obs_do if "obs_do" in emit else None, _abi.OS_NOBS, n, name="obs_do")
    poe, l4 = _dev_ptr(buf.obs_ec if "obs_ec" in emit else None, _abi.OS_NOBS, n, name="obs_ec")
    pse, l5 = _dev_ptr(buf.state if "state" in emit else None, _abi.OS_NSTATE, n, name="state")
    pdn, _ = _dev_ptr(buf.done, 1, n, dtype=torch.uint8, name="done")
    pss, _ = _dev_ptr(buf.status, 1, n, dtype=torch.int32, name="status")
    pct, l6 = _dev_ptr(buf.counters, 2, n, dtype=torch.int32, name="counters")
    pal = pol = None
    l7 = l8 = None
    if act_log is not None:
        if act_log.shape != (K, 2, n) or not act_log.is_contiguous():
            raise ValueError("act_log must be contiguous [K, 2, n]")
        pal, l7 = _dev_ptr(act_log.view(2 * K, n), 2 * K, n, name="act_log")
    if obs_log is not None:
        if obs_log.shape != (K, 2 * _abi.OS_NOBS, n) or not obs_log.is_contiguous():
            raise ValueError("obs_log must be contiguous [K, 18, n]")
        pol, l8 = _dev_ptr(obs_log.view(2 * _abi.OS_NOBS * K, n), 2 * _abi.OS_NOBS * K, n, name="obs_log")
    ld = _same_ld([l0, l1, l2, l3 if pod else None, l4 if poe else None, l5 if pse else None, l6, l7, l8], "os_rollout_k")
    tol = tol or _abi.make_tol()
    with torch.cuda.device(buf.st.device):
        rc = lib.sbr_os_rollout_k(n, ld, K, pst, pac, C.byref(policy), C.byref(params), C.byref(sched), pod, poe, pse,
                                  prw, pdn, pss, pct, pal, pol, int(mode), C.byref(tol), _stream_ptr(stream))
    _abi.check(rc, "sbr_os_rollout_k")
    return buf


class V4Buffers(object):
    """Device buffers of the SBR-v4 path: persistent state + per-step outputs (allocated once)."""

    def __init__(self, n, device):
        f = dict(dtype=torch.float64, device=device)
        self.st = torch.zeros((_abi.V4_ROWS, n), **f)
        self.obs = torch.empty((_abi.NX, n), **f)
        self.reward = torch.zeros((n,), **f)
        self.done = torch.ones((n,), dtype=torch.uint8, device=device)
        self.status = torch.zeros((n,), dtype=torch.int32, device=device)
        self.counters = torch.zeros((2, n), dtype=torch.int32, device=device)


def v4_reset(buf, influent, params, x0=None, mask=None, stream=None, order=None):
    """SbrEnv4.reset for a batch (gym_SBR_env4.py:94-198).  influent [14,n] (row 0 = fill flow).
    order: optional int32 [n] -- slot j of buf.st holds env order[j] (see sbr_v4_step in the header)."""
    lib = _abi.load()
    n = buf.st.shape[1]
    pst, l0 = _dev_ptr(buf.st, _abi.V4_ROWS, n, name="st")
    pin, l1 = _dev_ptr(influent, _abi.NX, n, name="influent")
    px0, l2 = _dev_ptr(x0, _abi.NX, n, name="x0")
    pmk, _ = _dev_ptr(mask, 1, n, dtype=torch.uint8, name="mask")
    pob, l3 = _dev_ptr(buf.obs, _abi.NX, n, name="obs")
    pdn, _ = _dev_ptr(buf.done, 1, n, dtype=torch.uint8, name="done")
    por, _ = _dev_ptr(order, 1, n, dtype=torch.int32, name="order")
    ld = _same_ld([l0, l1, l2 if x0 is not None else None, l3], "v4_reset")
    with torch.cuda.device(buf.st.device):
        rc = lib.sbr_v4_reset(n, ld, px0, pin, pmk, C.byref(params), pst, pob, pdn, por, _stream_ptr(stream))
    _abi.check(rc, "sbr_v4_reset")
    return buf


def v4_step(buf, influent, action, params, sched, mode=_abi.MODE_DP45, tol=None, stream=None, order=None):
    """SbrEnv4.step for a batch (gym_SBR_env4.py:200-358).  action [n]: change of the DO set-point.
    order: optional int32 [n] -- slot j of buf.st holds env order[j]; all other buffers are indexed by env."""
    lib = _abi.load()
    n = buf.st.shape[1]
    pst, l0 = _dev_ptr(buf.st, _abi.V4_ROWS, n, name="st")
    pin, l1 = _dev_ptr(influent, _abi.NX, n, name="influent")
    pac, _ = _dev_ptr(action, 1, n, name="action")
    pob, l2 = _dev_ptr(buf.obs, _abi.NX, n, name="obs")
    prw, _ = _dev_ptr(buf.reward, 1, n, name="reward")
    pdn, _ = _dev_ptr(buf.done, 1, n, dtype=torch.uint8, name="done")
    pss, _ = _dev_ptr(buf.status, 1, n, dtype=torch.int32, name="status")
    pct, l3 = _dev_ptr(buf.counters, 2, n, dtype=torch.int32, name="counters")
    por, _ = _dev_ptr(order, 1, n, dtype=torch.int32, name="order")
    ld = _same_ld([l0, l1, l2, l3], "v4_step")
    tol = tol or _abi.make_tol()
    with torch.cuda.device(buf.st.device):
        rc = lib.sbr_v4_step(n, ld, pst, pin, pac, C.byref(params), C.byref(sched), pob, prw, pdn, pss, pct,
                             int(mode), C.byref(tol), por, _stream_ptr(stream))
    _abi.check(rc, "sbr_v4_step")
    return buf


def v4_rollout_k(buf, influent, action, policy, rewards, params, sched, mode=_abi.MODE_DP45, tol=None, stream=None,
                 emit_obs=True, act_log=None, obs_log=None):
    """K = rewards.shape[0] consecutive SbrEnv4.step calls in ONE launch with the policy head (14 -> hidden -> 1) evaluated
    in-kernel between them (sbr_v4_rollout_k).  Every buffer is indexed like buf.st (no slot map).  action [n] in/out."""
    lib = _abi.load()
    n = buf.st.shape[1]
    K = rewards.shape[0]
    if rewards.shape != (K, n) or not rewards.is_contiguous():
        raise ValueError("rewards must be contiguous [K, n]")
    pst, l0 = _dev_ptr(buf.st, _abi.V4_ROWS, n, name="st")
    pin, l1 = _dev_ptr(influent, _abi.NX, n, name="influent")
    pac, _ = _dev_ptr(action, 1, n, name="action")
    prw, l2 = _dev_ptr(rewards, K, n, name="rewards") if K > 1 else (C.c_void_p(rewards.data_ptr()), None)
    pob, l3 = _dev_ptr(buf.obs if emit_obs else None, _abi.NX, n, name="obs")
    pdn, _ = _dev_ptr(buf.done, 1, n, dtype=torch.uint8, name="done")
    pss, _ = _dev_ptr(buf.status, 1, n, dtype=torch.int32, name="status")
    pct, l4 = _dev_ptr(buf.counters, 2, n, dtype=torch.int32, name="counters")
    pal = pol = None
    l5 = l6 = None
    if act_log is not None:
        if act_log.shape != (K, n) or not act_log.is_contiguous():
            raise ValueError("act_log must be contiguous [K, n]")
        pal, l5 = _dev_ptr(act_log, K, n, name="act_log") if K > 1 else (C.c_void_p(act_log.data_ptr()), None)
    if obs_log is not None:
        if obs_log.shape != (K, _abi.NX, n) or not obs_log.is_contiguous():
            raise ValueError("obs_log must be contiguous [K, 14, n]")
        pol, l6 = _dev_ptr(obs_log.view(_abi.NX * K, n), _abi.NX * K, n, name="obs_log")
    ld = _same_ld([l0, l1, l2, l3 if emit_obs else None, l4, l5, l6], "v4_rollout_k")
    tol = tol or _abi.make_tol()
    with torch.cuda.device(buf.st.device):
        rc = lib.sbr_v4_rollout_k(n, ld, K, pst, pin, pac, C.byref(policy), C.byref(params), C.byref(sched), pob, prw, pdn,
                                  pss, pct, pal, pol, int(mode), C.byref(tol), _stream_ptr(stream))
    _abi.check(rc, "sbr_v4_rollout_k")
    return buf


_INFLUENT_TABLES = {}


def influent_mix(switch, rnd, out=None, stream=None):
    """influent_mixed [14,n] from rnd [48,n] (standard-normal draws) for scenario `switch` -- the device version of
    buffer_tank3.influent.buffer_tank (buffer_tank3.py:18-108), bit-identical to numpy for the same rnd."""
    from . import influent as influent_mod
    lib = _abi.load()
    n = rnd.shape[1]
    key = (int(switch), rnd.device)
    if key not in _INFLUENT_TABLES:
        t = influent_mod.tables()
        mean = torch.as_tensor(t["mean"][int(switch)], dtype=torch.float64).contiguous()
        std = torch.as_tensor(t["std_frac"][int(switch)][:, None] * t["mean"][int(switch)], dtype=torch.float64).contiguous()
        _INFLUENT_TABLES[key] = (mean.to(rnd.device), std.to(rnd.device))
    mean, std = _INFLUENT_TABLES[key]
    if out is None:
        out = torch.empty((_abi.NX, n), dtype=torch.float64, device=rnd.device)
    pr, l0 = _dev_ptr(rnd, influent_mod.N_POINTS, n, name="rnd")
    po, l1 = _dev_ptr(out, _abi.NX, n, name="influent")
    ld = _same_ld([l0, l1], "influent_mix")
    with torch.cuda.device(rnd.device):
        rc = lib.sbr_influent_mix(n, ld, pr, C.c_void_p(mean.data_ptr()), C.c_void_p(std.data_ptr()), po,
                                  _stream_ptr(stream))
    _abi.check(rc, "sbr_influent_mix")
    return out


_ALL_TABLES = {}


def _all_tables(device, table_set="buffer_tank3"):
    """Device copies of all 8 scenario tables: mean, std [8,14,48].  table_set="buffer_tank2": the one live branch of
    buffer_tank2 (`SBR-v0/1`) in slot 0 -- same mean + std * rnd form, so the same sampler kernel serves it."""
    from . import influent as influent_mod
    key = (torch.device(device), table_set)
    if key not in _ALL_TABLES:
        if table_set == "buffer_tank2":
            m2, s2 = influent_mod.tables_bt2()
            mean = torch.zeros((8, _abi.NX, influent_mod.N_POINTS), dtype=torch.float64)
            std = torch.zeros_like(mean)
            mean[0], std[0] = torch.as_tensor(m2), torch.as_tensor(s2)
        elif table_set == "buffer_tank3":
            t = influent_mod.tables()
            mean = torch.as_tensor(t["mean"], dtype=torch.float64).contiguous()
            std = torch.as_tensor(t["std_frac"][:, :, None] * t["mean"], dtype=torch.float64).contiguous()
        else:
            raise ValueError("unknown influent table set %r" % (table_set,))
        _ALL_TABLES[key] = (mean.to(key[0]), std.to(key[0]))
    return _ALL_TABLES[key]


def influent_sample(n, device, seed, env_offset=0, scenario=0, epoch=None, epoch0=0, mask=None, out=None,
                    scenario_out=None, stream=None, table_set="buffer_tank3"):
    """N independent buffer_tank(scenario) calls (buffer_tank3.py:18-108) with counter-based randomness: the draws
    of env i depend only on (seed, env_offset + i, its episode number) -- see sbr_influent_sample in the header.
    scenario: 0..7 or -1 (drawn per env, SbrEnv4.reset); epoch: optional int64 [n], incremented for drawing envs."""
    lib = _abi.load()
    mean, std = _all_tables(device, table_set)
    if table_set == "buffer_tank2" and int(scenario) != 0:
        raise ValueError("buffer_tank2 has one live scenario (slot 0)")
    if out is None:
        out = torch.zeros((_abi.NX, n), dtype=torch.float64, device=device)
    po, ld = _dev_ptr(out, _abi.NX, n, name="influent")
    pe, _ = _dev_ptr(epoch, 1, n, dtype=torch.int64, name="epoch")
    pm, _ = _dev_ptr(mask, 1, n, dtype=torch.uint8, name="mask")
    ps, _ = _dev_ptr(scenario_out, 1, n, dtype=torch.int32, name="scenario_out")
    with torch.cuda.device(device):
        rc = lib.sbr_influent_sample(n, ld, int(seed) & 0xFFFFFFFFFFFFFFFF, int(env_offset), pe, int(epoch0),
                                     int(scenario), C.c_void_p(mean.data_ptr()), C.c_void_p(std.data_ptr()), pm, po,
                                     ps, _stream_ptr(stream))
    _abi.check(rc, "sbr_influent_sample")
    return out


def philox_normals(n, device, seed, env_offset=0, epoch0=0, stream=None):
    """The 48 standard normals per env that influent_sample draws for episode `epoch0`: z [48,n]."""
    from . import influent as influent_mod
    lib = _abi.load()
    z = torch.empty((influent_mod.N_POINTS, n), dtype=torch.float64, device=device)
    pz, ld = _dev_ptr(z, influent_mod.N_POINTS, n, name="z")
    with torch.cuda.device(device):
        rc = lib.sbr_philox_normals(n, ld, int(seed) & 0xFFFFFFFFFFFFFFFF, int(env_offset), int(epoch0), pz,
                                    _stream_ptr(stream))
    _abi.check(rc, "sbr_philox_normals")
    return z


def permute_rows(perm, pairs, scatter=False, stream=None):
    """Row permutation of SoA buffers in one launch: for every (src, dst) pair of [rows, n] (or [n]) tensors,
    dst[r, i] = src[r, perm[i]] (gather) or dst[r, perm[i]] = src[r, i] (scatter).  float64 / int32 / uint32."""
    lib = _abi.load()
    n = perm.shape[0]
    if not (0 < len(pairs) <= _abi.PERMUTE_MAX):
        raise ValueError("permute_rows: 1..%d buffers per launch" % _abi.PERMUTE_MAX)
    pp, _ = _dev_ptr(perm, 1, n, dtype=torch.int64, name="perm")
    k = len(pairs)
    src, dst = (C.c_void_p * k)(), (C.c_void_p * k)()
    lds, ldd = (C.c_int64 * k)(), (C.c_int64 * k)()
    rows, elem = (C.c_int32 * k)(), (C.c_int32 * k)()
    for j, (s, d) in enumerate(pairs):
        if s.dtype != d.dtype or s.shape != d.shape or not (s.is_cuda and d.is_cuda):
            raise ValueError("permute_rows: src/dst of pair %d differ in dtype, shape or device" % j)
        if s.element_size() not in (4, 8):
            raise TypeError("permute_rows: 4- or 8-byte elements only")
        if s.dim() == 1:
            s, d = s[None, :], d[None, :]
        if s.shape[1] != n or s.stride(1) != 1 or d.stride(1) != 1:
            raise ValueError("permute_rows: buffers must be [rows, %d] with unit stride along envs" % n)
        src[j], dst[j] = s.data_ptr(), d.data_ptr()
        lds[j] = s.stride(0) if s.shape[0] > 1 else n
        ldd[j] = d.stride(0) if d.shape[0] > 1 else n
        rows[j], elem[j] = s.shape[0], s.element_size()
    with torch.cuda.device(perm.device):
        rc = lib.sbr_permute_rows(n, pp, k, src, dst, lds, ldd, rows, elem, 1 if scatter else 0, _stream_ptr(stream))
    _abi.check(rc, "sbr_permute_rows")


def policy_mlp(obs_a, obs_b, w1, w2, lo, span, action, stream=None):
    """action[o] = lo[o] + span[o] * sigmoid(w2 @ tanh(w1 @ [obs_a; obs_b])) per env in one launch (sbr_policy_mlp).
    obs_a [ra,n], obs_b [rb,n] (or None) float64 SoA; w1 [hidden, ra+rb], w2 [n_out, hidden], lo / span [n_out] float32;
    action [n_out, n] float64 out."""
    lib = _abi.load()
    n = obs_a.shape[1]
    ra, rb = obs_a.shape[0], (0 if obs_b is None else obs_b.shape[0])
    hidden, n_out = w1.shape[0], w2.shape[0]
    pa, l0 = _dev_ptr(obs_a, ra, n, name="obs_a")
    pb, l1 = _dev_ptr(obs_b, rb, n, name="obs_b") if obs_b is not None else (None, None)
    pact, l2 = _dev_ptr(action, n_out, n, name="action") if n_out > 1 else (_dev_ptr(action.reshape(-1), 1, n)[0], None)
    ld = _same_ld([l0, l1, l2], "policy_mlp")
    for t, shape, name in ((w1, (hidden, ra + rb), "w1"), (w2, (n_out, hidden), "w2"), (lo, (n_out,), "lo"),
                           (span, (n_out,), "span")):
        if not t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != shape or not t.is_contiguous():
            raise ValueError("%s must be a contiguous CUDA float32 tensor of shape %s" % (name, shape))
    with torch.cuda.device(obs_a.device):
        rc = lib.sbr_policy_mlp(n, ld, pa, ra, pb, rb, C.c_void_p(w1.data_ptr()), C.c_void_p(w2.data_ptr()),
                                C.c_void_p(lo.data_ptr()), C.c_void_p(span.data_ptr()), hidden, n_out, pact,
                                _stream_ptr(stream))
    _abi.check(rc, "sbr_policy_mlp")
    return action


def reward_stats(reward, status=None, out=None, stream=None):
    """[sum, sumsq, min, max, count] of the rewards of healthy envs, on the device (feeds the NCCL gather)."""
    lib = _abi.load()
    n = reward.shape[0]
    if out is None:
        out = torch.empty((5,), dtype=torch.float64, device=reward.device)
    pr, _ = _dev_ptr(reward, 1, n, name="reward")
    ps, _ = _dev_ptr(status, 1, n, dtype=torch.int32, name="status")
    sp = _stream_ptr(stream)
    with torch.cuda.device(reward.device):
        _abi.check(lib.sbr_reward_stats_init(C.c_void_p(out.data_ptr()), sp), "sbr_reward_stats_init")
        _abi.check(lib.sbr_reward_stats(n, pr, ps, C.c_void_p(out.data_ptr()), sp), "sbr_reward_stats")
    return out


def fp64_probe(blocks, threads, iters, device, stream=None):
    """Launch the DFMA probe once; returns (sink tensor, flops issued).  Time it with CUDA events."""
    lib = _abi.load()
    sink = torch.empty((blocks * threads,), dtype=torch.float64, device=device)
    flops = C.c_double(0.0)
    with torch.cuda.device(device):
        rc = lib.sbr_fp64_probe(int(blocks), int(threads), int(iters), C.c_void_p(sink.data_ptr()), C.byref(flops),
                                _stream_ptr(stream))
    _abi.check(rc, "sbr_fp64_probe")
    return sink, flops.value
