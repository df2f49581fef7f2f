"""ctypes binding of the CPU twin (TEST INFRASTRUCTURE ONLY; see sbr_twin.cpp).  numpy in / numpy out, SoA."""
import ctypes as C
import os
import subprocess

import numpy as np

from gym_sbr2_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(os.path.dirname(HERE), "_build")
LIB = os.path.join(BUILD_DIR, "libsbr_twin.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "sbr_twin.cpp")
    csrc = os.path.join(os.path.dirname(os.path.dirname(HERE)), "gym_sbr2_b200", "csrc")
    deps = [src, os.path.join(csrc, "sbr_core.cuh"), os.path.join(csrc, "sbr_cnt.cuh"), os.path.join(csrc, "sbr_ilc.cuh"),
            os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "sbr_b200.h")]
    os.makedirs(BUILD_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, src])
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
    return _lib


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def default_params():
    """Defaults without touching the CUDA library (mirrors sbr_params_default in sbr_kernels.cu)."""
    p = _abi.SbrParams()
    vals = dict(muh=4.0, Ks=10.0, Koh=0.2, Kno=0.5, bh=0.3, etag=0.8, etah=0.8, kh=3.0, Kx=0.1, mua=0.5, Knh=1.0,
                ba=0.05, Koa=0.4, ka=0.05, Ya=0.24, Yh=0.67, fp=0.08, ixb=0.08, ixp=0.06,
                pid_Kc=5.0, pid_tauI=0.00035, pid_tauD=0.005, pid_dt=0.02 / 24, kla_min=0.0, kla_max=240.0,
                WV=1.32, Qin=1.32 - 0.6161484733495801, Qeff=0.66, biomass_setpoint=2700.0,
                settler_area=(1.25 / 2) * (1.25 / 2), settler_vmax=474.0, kla0=0.0, action_scale=8.0,
                os_Kc_DO=100.0, os_tauI_DO=20.0, os_tauD_DO=0.0, os_Kc_EC=100.0, os_tauI_EC=20.0, os_tauD_EC=0.0,
                os_pid_dt=0.002 / 24, ec_min=0.0, ec_max=0.0005, ec_conc=1200000.0 * 4, do_sp_max=8.0,
                no_sp_max=15.0, IV=0.6161484733495801)
    tk = (15 + 273.15) / 100
    vals["so_sat"] = 0.9997743214 * (8 / 10.5) * 6791.5 * (56.12 * np.exp(-66.7354 + 87.4755 / tk + 24.4526 * np.log(tk)))
    for k, v in vals.items():
        setattr(p, k, float(v))
    return p


def cycle_v2(x0, influent, action, params, sched, mode=0, tol=None):
    """x0, influent: [14, n]; action: [3, n] float64 C-contiguous.  Returns dict of numpy arrays."""
    lib = load()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    influent = np.ascontiguousarray(influent, dtype=np.float64)
    action = np.ascontiguousarray(action, dtype=np.float64)
    n = x0.shape[1]
    x_last = np.empty((14, n)); obs = np.empty((3, n)); reward = np.empty(n); aux = np.empty((_abi.AUX_ROWS, n))
    status = np.zeros(n, dtype=np.int32); counters = np.zeros((2, n), dtype=np.uint32)
    tol = tol or _abi.make_tol()
    rc = lib.twin_cycle_v2(C.c_int64(n), C.c_int64(n), _ptr(x0), _ptr(influent), _ptr(action), C.byref(params),
                           C.byref(sched), _ptr(x_last), _ptr(obs), _ptr(reward), _ptr(aux), _ptr(status),
                           _ptr(counters), C.c_int(mode), C.byref(tol))
    assert rc == 0
    return dict(x_last=x_last, obs=obs, reward=reward, aux=aux, status=status, counters=counters)


def integrate_interval(x, kla, params, tail, T, n_sub, mode=0, tol=None, ec=None, loading=None):
    lib = load()
    x = np.array(x, dtype=np.float64, order="C")
    n = x.shape[1]
    kla = np.ascontiguousarray(kla, dtype=np.float64)
    ec = None if ec is None else np.ascontiguousarray(ec, dtype=np.float64)
    loading = None if loading is None else np.ascontiguousarray(loading, dtype=np.float64)
    counters = np.zeros((2, n), dtype=np.uint32)
    tol = tol or _abi.make_tol()
    rc = lib.twin_integrate_interval(C.c_int64(n), C.c_int64(n), _ptr(x), _ptr(kla), _ptr(ec), _ptr(loading),
                                     C.byref(params), C.c_int(tail), C.c_double(T), C.c_int(n_sub), C.c_int(mode),
                                     C.byref(tol), _ptr(counters))
    assert rc == 0
    return x, counters


def rhs(x, kla, params, tail, ec=None, loading=None):
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[1]
    kla = np.ascontiguousarray(kla, dtype=np.float64)
    ec = None if ec is None else np.ascontiguousarray(ec, dtype=np.float64)
    loading = None if loading is None else np.ascontiguousarray(loading, dtype=np.float64)
    dx = np.zeros((14, n))
    rc = lib.twin_rhs(C.c_int64(n), C.c_int64(n), _ptr(x), _ptr(kla), _ptr(ec), _ptr(loading), C.byref(params),
                      C.c_int(tail), _ptr(dx))
    assert rc == 0
    return dx


class OsBatch(object):
    """Host-memory twin of the interval-per-step path: same buffers and call sequence as the CUDA entry points
    sbr_os_reset / sbr_os_step."""

    def __init__(self, n, params=None, sched=None, mode=0, tol=None):
        from gym_sbr2_b200 import schedule
        self.n = n
        self.params = params or default_params()
        self.sched = sched or schedule.os_schedule()
        self.mode = mode
        self.tol = tol or _abi.make_tol()
        self.st = np.zeros((_abi.OS_ROWS, n))
        self.obs_do = np.zeros((9, n)); self.obs_ec = np.zeros((9, n)); self.state = np.zeros((15, n))
        self.reward = np.zeros(n); self.done = np.zeros(n, dtype=np.uint8)
        self.status = np.zeros(n, dtype=np.int32); self.counters = np.zeros((2, n), dtype=np.uint32)

    def reset(self, influent, x0=None, mask=None):
        lib = load()
        influent = np.ascontiguousarray(influent, dtype=np.float64)
        x0 = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        n = self.n
        rc = lib.twin_os_reset(C.c_int64(n), C.c_int64(n), _ptr(x0), _ptr(influent), _ptr(mask),
                               C.byref(self.params), C.byref(self.sched), _ptr(self.st), _ptr(self.obs_do),
                               _ptr(self.obs_ec), _ptr(self.done), _ptr(self.status), _ptr(self.counters),
                               C.c_int(self.mode), C.byref(self.tol))
        assert rc == 0
        return self.obs_do.copy(), self.obs_ec.copy()

    def step(self, action):
        lib = load()
        action = np.ascontiguousarray(action, dtype=np.float64)
        n = self.n
        assert action.shape == (2, n)
        rc = lib.twin_os_step(C.c_int64(n), C.c_int64(n), _ptr(self.st), _ptr(action), C.byref(self.params),
                              C.byref(self.sched), _ptr(self.obs_do), _ptr(self.obs_ec), _ptr(self.state),
                              _ptr(self.reward), _ptr(self.done), _ptr(self.status), _ptr(self.counters),
                              C.c_int(self.mode), C.byref(self.tol))
        assert rc == 0
        return self.obs_do.copy(), self.obs_ec.copy(), self.state.copy(), self.reward.copy(), self.done.copy()


class V4Batch(object):
    """Host-memory twin of the SBR-v4 path: same buffers and call sequence as sbr_v4_reset / sbr_v4_step."""

    def __init__(self, n, params=None, sched=None, mode=0, tol=None):
        from gym_sbr2_b200 import schedule
        self.n = n
        self.params = params or default_params()
        self.sched = sched or schedule.os_schedule()
        self.mode = mode
        self.tol = tol or _abi.make_tol()
        self.st = np.zeros((_abi.V4_ROWS, n))
        self.obs = np.zeros((14, n)); self.reward = np.zeros(n); self.done = np.ones(n, dtype=np.uint8)
        self.status = np.zeros(n, dtype=np.int32); self.counters = np.zeros((2, n), dtype=np.uint32)
        self.influent = np.zeros((14, n))

    def reset(self, influent, x0=None, mask=None):
        lib = load()
        self.influent = np.ascontiguousarray(influent, dtype=np.float64)
        x0 = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        n = self.n
        rc = lib.twin_v4_reset(C.c_int64(n), C.c_int64(n), _ptr(x0), _ptr(self.influent), _ptr(mask),
                               C.byref(self.params), _ptr(self.st), _ptr(self.obs), _ptr(self.done))
        assert rc == 0
        return self.obs.copy()

    def step(self, action):
        lib = load()
        action = np.ascontiguousarray(action, dtype=np.float64).reshape(-1)
        n = self.n
        assert action.shape == (n,)
        rc = lib.twin_v4_step(C.c_int64(n), C.c_int64(n), _ptr(self.st), _ptr(self.influent), _ptr(action),
                              C.byref(self.params), C.byref(self.sched), _ptr(self.obs), _ptr(self.reward),
                              _ptr(self.done), _ptr(self.status), _ptr(self.counters), C.c_int(self.mode),
                              C.byref(self.tol))
        assert rc == 0
        return self.obs.copy(), self.reward.copy(), self.done.copy()


class CntBatch(object):
    """Host-memory twin of the SBRCnt-v0/1/2, SBRCntMA-v1, SBROS-v2 path: same buffers and call sequence as
    sbr_cnt_reset / sbr_cnt_step.  `cfg`: an _abi.SbrCntConfig (gym_sbr2_b200.cnt.cnt_config(kind))."""

    def __init__(self, cfg, n, obs_rows, params=None, sched=None, mode=1, tol=None):
        from gym_sbr2_b200 import schedule
        self.cfg = cfg
        self.n = n
        self.params = params or default_params()
        self.sched = sched or schedule.os_schedule()
        self.mode = mode
        self.tol = tol or _abi.make_tol()
        self.st = np.zeros((_abi.CNT_ROWS, n))
        self.obs = np.zeros((obs_rows, n)); self.reward = np.zeros(n); self.done = np.ones(n, dtype=np.uint8)
        self.status = np.zeros(n, dtype=np.int32); self.counters = np.zeros((2, n), dtype=np.uint32)

    def reset(self, influent, x0=None, mask=None):
        lib = load()
        influent = np.ascontiguousarray(influent, dtype=np.float64)
        x0 = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        n = self.n
        rc = lib.twin_cnt_reset(C.c_int64(n), C.c_int64(n), C.byref(self.cfg), _ptr(x0), _ptr(influent), _ptr(mask),
                                C.byref(self.params), C.byref(self.sched), _ptr(self.st), _ptr(self.obs),
                                _ptr(self.done), _ptr(self.status), _ptr(self.counters), C.c_int(self.mode),
                                C.byref(self.tol))
        assert rc == 0
        return self.obs.copy()

    def step(self, action):
        lib = load()
        action = np.ascontiguousarray(action, dtype=np.float64)
        n = self.n
        assert action.shape == (2, n)
        rc = lib.twin_cnt_step(C.c_int64(n), C.c_int64(n), C.byref(self.cfg), _ptr(self.st), _ptr(action),
                               C.byref(self.params), C.byref(self.sched), _ptr(self.obs), _ptr(self.reward),
                               _ptr(self.done), _ptr(self.status), _ptr(self.counters), C.c_int(self.mode),
                               C.byref(self.tol))
        assert rc == 0
        return self.obs.copy(), self.reward.copy(), self.done.copy()


def cycle_ilc(x0, influent, sp, params, sched, layout, t_fill, kla_base=None, u=None, want_kla_mem=True, mode=1,
              tol=None, want_so_mem=True):
    """x0, influent [14, n]; sp [3, n]; kla_base, u [S, n] or None (cycle 0).  Returns dict of numpy arrays."""
    lib = load()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    influent = np.ascontiguousarray(influent, dtype=np.float64)
    sp = np.ascontiguousarray(sp, dtype=np.float64)
    n, S = x0.shape[1], int(layout.n_samples)
    kla_base = None if kla_base is None else np.ascontiguousarray(kla_base, dtype=np.float64)
    u = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
    so_mem = np.zeros((S, n)) if want_so_mem else None; kla_mem = np.zeros((S, n)) if want_kla_mem else None
    x_last = np.empty((14, n)); out = np.empty((_abi.ILC_OUT_ROWS, n)); status = np.zeros(n, dtype=np.int32)
    counters = np.zeros((2, n), dtype=np.uint32)
    tol = tol or _abi.make_tol(1e-9, 1e-11)         # gym_sbr2_b200.ilc.ILC_RTOL / ILC_ATOL
    lib.twin_cycle_ilc.restype = C.c_int
    rc = lib.twin_cycle_ilc(C.c_int64(n), C.c_int64(n), _ptr(x0), _ptr(influent), _ptr(sp), C.byref(params),
                            C.byref(sched), C.byref(layout), C.c_double(t_fill), _ptr(kla_base), _ptr(u), _ptr(so_mem),
                            _ptr(kla_mem), _ptr(x_last), _ptr(out), _ptr(status), _ptr(counters), C.c_int(mode),
                            C.byref(tol))
    assert rc == 0
    return dict(x_last=x_last, so_mem=so_mem, kla_mem=kla_mem, out=out, status=status, counters=counters)


def ilc_update(layout, w, D, sp6, so_mem, e_sum, e_last, dt, Kc, tauI, tauD):
    """e_sum, e_last [S, n] are updated in place; returns u [S, n]."""
    lib = load()
    n = so_mem.shape[1]
    w = np.ascontiguousarray(w, dtype=np.float64); D = np.ascontiguousarray(D, dtype=np.float64)
    sp6 = np.ascontiguousarray(sp6, dtype=np.float64); so_mem = np.ascontiguousarray(so_mem, dtype=np.float64)
    assert e_sum.flags.c_contiguous and e_last.flags.c_contiguous
    u = np.zeros_like(so_mem)
    rc = lib.twin_ilc_update(C.c_int64(n), C.c_int64(n), C.byref(layout), _ptr(w), _ptr(D), _ptr(sp6), _ptr(so_mem),
                             _ptr(e_sum), _ptr(e_last), _ptr(u), C.c_double(dt), C.c_double(Kc), C.c_double(tauI),
                             C.c_double(tauD))
    assert rc == 0
    return u


def cycle_v2_traj(x0, influent, action, params, sched, t_start, mode=0, tol=None):
    """The cycle with its trajectory record: returns dict(x_last, obs, reward, traj [R, 16, n])."""
    lib = load()
    x0 = np.ascontiguousarray(x0, dtype=np.float64); influent = np.ascontiguousarray(influent, dtype=np.float64)
    action = np.ascontiguousarray(action, dtype=np.float64)
    n = x0.shape[1]
    R = 1 + sum(sched.n_int[k] for k in range(8) if k not in (5, 6))
    traj = np.full((R, _abi.TRAJ2_ROWS, n), np.nan)
    x_last = np.empty((14, n)); obs = np.empty((3, n)); reward = np.empty(n)
    ts = (C.c_double * 8)(*[float(v) for v in t_start])
    tol = tol or _abi.make_tol()
    rc = lib.twin_cycle_v2_traj(C.c_int64(n), C.c_int64(n), _ptr(x0), _ptr(influent), _ptr(action), C.byref(params),
                                C.byref(sched), ts, _ptr(x_last), _ptr(obs), _ptr(reward), _ptr(traj), C.c_int(mode),
                                C.byref(tol))
    assert rc == 0
    return dict(x_last=x_last, obs=obs, reward=reward, traj=traj)
