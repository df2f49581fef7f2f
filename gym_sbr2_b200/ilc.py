"""Batch-to-batch (iterative-learning) feed-forward KLa of the reference's `SBR-v0` on the CUDA path.

Reference: gym_SBR_env0.py (module state + SbrEnv), module_batch_PID.py:7-275 (batch_PID), SBR_model_PID_on.py /
sub_phases_PID_on.py (cycle 0), SBR_model_batchPID_fbPID.py / sub_phases_batchPID_fbPID.py (feed-forward cycles).
`SbrEnv.step` cannot run in the reference (float `num` in np.linspace; a seven-argument call of the ten-parameter
module_reward.sbr_reward, gym_SBR_env0.py:203), so what is reproduced -- and pinned by tests against the reference's own
functions -- is everything `step()` does before that call: `_take_action` (batch_PID) and `_next_observation` (the cycle).
The reward returned here is module_reward.sbr_reward's formula on this cycle's applied KLa: by construction, not pinned.
The influent source of these two envs, buffer_tank2.influent.buffer_tank(0, 12), is gym_sbr2_b200/influent.py's
`mix_numpy_bt2` (bit-exact) and, on the device, the counter-based sampler with the buffer_tank2 tables.

Host side of the path: the constant tables of the batch-to-batch controller (window weights and their sums, computed
with the reference's own expressions, mix-ups included), the sample layout, the torch-facing wrappers of the two C
entry points, and `SbrIlcVecEnv`.  All arithmetic on the env state happens in libsbr_b200.so; there is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi, core, schedule

# gym_SBR_env0.py:40-41,69-71,76,89,93
WV, IV = 1.32, 0.66
X0_ILC = [IV, 30.0, 0.5601630529230822, 1762.3890076468106, 30.97046860269441, 2628.6551849696393, 188.71238190722482,
          780.479571994941, 6.83620016588177, 14.575400491942467, 0.00872090237410032, 0.36940333660700486,
          1.896711744868243, 3.705237172170034]
FILL_FLOW = 31.4285
PID = dict(Kc=0.5 / 1.18, tauI=0.0015, tauD=0.005, dt=0.05, lo=0.0, hi=240.0)
PAR_BATCH_PID = [0.002018, 0.003643, 0.004036, 0, 0.01875, 0.0004671, 0.01564, 0.003643, 0.001028, 0, 0, 0, 0, 0,
                 0.003027, 0.003643]
KC_B, TAUI_B, TAUD_B = 1 / 1.18, 0.25, 0.1           # module_batch_PID.py:15-17
BIOMASS_SETPOINT = 5400.0                             # SBR_model_batchPID_fbPID.py:283
CYCLE0_SETPOINTS = (2.0, 2.0, 2.0)                    # gym_SBR_env0.py:98: DO_setpoints = [0,0,2,0,2,0,0,2]
ACTION_LOW, ACTION_HIGH = 0.0, 5.0                    # gym_SBR_env0.py:145
OBS_SCALE = [1.0, 60, 31, 1974, 107, 2237, 195, 988, 2, 4, 14, 3, 5, 12]   # gym_SBR_env0.py:159-172 (state[0] := 1)
PHASES = (0, 1, 2, 3, 4, 7)                           # schedule index of the six PID-controlled phases
# Default tolerance of the adaptive cycle: the So memory comes from the stepper's 4th-order continuous extension, whose
# error is ~10x the step's.  Measured on cycle 0 against LSODA at 1e-12 (tests/test_twin_parity_ilc.py): rtol 1e-8 ->
# 1.4e-7 g/m3 with 8.6 k right-hand sides, 1e-9 -> 1.0e-8 g/m3 with 11.2 k; the reference's own memory is 6.3e-7 off.
ILC_RTOL, ILC_ATOL = 1e-9, 1e-11


def apply_constants(p):
    """Overwrite the controller / plant fields of an SbrParams with this path's constants (in place)."""
    p.pid_Kc, p.pid_tauI, p.pid_tauD, p.pid_dt = PID["Kc"], PID["tauI"], PID["tauD"], PID["dt"]
    p.kla_min, p.kla_max = PID["lo"], PID["hi"]
    p.WV, p.IV, p.Qin = WV, IV, WV - IV
    p.biomass_setpoint = BIOMASS_SETPOINT
    return p


def layout(sched, par=PAR_BATCH_PID, t_delta=schedule.T_DELTA):
    """SbrIlcLayout of a cycle schedule on the reference's output grid: offsets of the six phases in the [S][N] sample
    memories and the window lengths tp = int(3 tau_w / t_delta) (module_batch_PID.py:29)."""
    lay = _abi.SbrIlcLayout()
    off = 0
    for j, k in enumerate(PHASES):
        lay.off[j] = off
        off += sched.n_int[k] * sched.n_sub[k] + 1
        lay.tp[j] = int(par[2 * k] * 3 / t_delta)
    lay.n_samples = off
    return lay


def weights(sched, par=PAR_BATCH_PID, t_delta=schedule.T_DELTA):
    """(w [S], D [S]) float64: window weights of the batch-to-batch error and their window sums D(t) = sum w dt over
    [t, min(t + tp, n)) (module_batch_PID.py:20-52).  w(t) = ((t - theta)/tau_a) exp(-(t - theta_b)/tau_b) above theta,
    0 below -- with the reference's own parameter mix-ups: phases 2-4 divide the linear factor by tau_w1, phase 3's exponent
    uses theta_w1 and tau_w1 (:66,94,122)."""
    stamps = schedule.phase_stamps(t_delta=t_delta)
    lay = layout(sched, par, t_delta)
    w_all, d_all = [], []
    tau1, theta1 = par[0], par[1]
    for j, k in enumerate(PHASES):
        t = np.asarray(stamps[k], dtype=np.float64)
        n = sched.n_int[k] * sched.n_sub[k] + 1
        if len(t) != n:
            raise ValueError("phase %d: %d stamps for %d samples (the schedule must be the reference grid)" % (k + 1, len(t), n))
        tau, theta = par[2 * k], par[2 * k + 1]
        idx = int(np.where(t > theta)[0][0])
        ts = t[idx:]
        if j in (0, 4, 5):
            w2 = ((ts - theta) / tau) * np.exp(-((ts - theta) / tau))
        elif j == 2:
            w2 = ((ts - theta) / tau1) * np.exp(-((ts - theta1) / tau1))
        else:
            w2 = ((ts - theta) / tau1) * np.exp(-((ts - theta) / tau))
        w = np.concatenate([np.zeros(idx), w2])
        tp = lay.tp[j]
        wd = np.concatenate([w * t_delta, np.zeros(tp)])
        d = np.lib.stride_tricks.sliding_window_view(wd, tp)[:n].sum(axis=1) if tp > 0 else np.zeros(n)
        w_all.append(w); d_all.append(d)
    return np.concatenate(w_all), np.concatenate(d_all), lay


class IlcCycleOut(object):
    """Output buffers of sbr_cycle_ilc.  so_mem / kla_mem may be handed in (the env points them at its own memories so that
    nothing is copied between the launches); kla_mem=False: the [S][n] KLa profile is not written at all."""

    def __init__(self, n, n_samples, device, so_mem=None, kla_mem=None):
        f = dict(dtype=torch.float64, device=device)
        self.so_mem = None if so_mem is False else (torch.zeros((n_samples, n), **f) if so_mem is None else so_mem)
        self.kla_mem = None if kla_mem is False else (torch.zeros((n_samples, n), **f) if kla_mem is None else kla_mem)
        self.x_last = torch.empty((_abi.NX, n), **f)
        self.out = torch.empty((_abi.ILC_OUT_ROWS, n), **f)
        self.status = torch.empty((n,), dtype=torch.int32, device=device)
        self.counters = torch.empty((2, n), dtype=torch.int32, device=device)


def cycle_ilc(x0, influent, sp, params, sched, lay, kla_base=None, u=None, out=None, t_fill=None, mode=_abi.MODE_DP45,
              tol=None, stream=None):
    """sbr_cycle_ilc.  x0, influent [14,n]; sp [3,n]; kla_base, u [S,n] (both None = cycle 0)."""
    lib = _abi.load()
    n, S = x0.shape[1], int(lay.n_samples)
    if out is None:
        out = IlcCycleOut(n, S, x0.device)
    px0, l0 = core._dev_ptr(x0, _abi.NX, n, name="x0")
    pin, l1 = core._dev_ptr(influent, _abi.NX, n, name="influent")
    psp, l2 = core._dev_ptr(sp, 3, n, name="sp")
    pkb, l3 = core._dev_ptr(kla_base, S, n, name="kla_base")
    pu, l4 = core._dev_ptr(u, S, n, name="u")
    pso, l5 = core._dev_ptr(out.so_mem, S, n, name="so_mem")
    pkm, l6 = core._dev_ptr(out.kla_mem, S, n, name="kla_mem")
    pxl, l7 = core._dev_ptr(out.x_last, _abi.NX, n, name="x_last")
    pout, l8 = core._dev_ptr(out.out, _abi.ILC_OUT_ROWS, n, name="out")
    pst, _ = core._dev_ptr(out.status, 1, n, dtype=torch.int32, name="status")
    pct, l9 = core._dev_ptr(out.counters, 2, n, dtype=torch.int32, name="counters")
    tol = tol or _abi.make_tol(ILC_RTOL, ILC_ATOL)
    lds = [l0, l1, l2, l7, l8, l9] + ([l3, l4] if kla_base is not None else []) + ([l5] if out.so_mem is not None else []) \
        + ([l6] if out.kla_mem is not None else [])
    ld = core._same_ld(lds, "cycle_ilc")
    t_fill = schedule.T_CYCLE * schedule.T_RATIO[0] if t_fill is None else float(t_fill)
    with torch.cuda.device(x0.device):
        rc = lib.sbr_cycle_ilc(n, ld, px0, pin, psp, C.byref(params), C.byref(sched), C.byref(lay), t_fill, pkb, pu, pso,
                               pkm, pxl, pout, pst, pct, int(mode), C.byref(tol), core._stream_ptr(stream))
    _abi.check(rc, "sbr_cycle_ilc")
    return out


def ilc_update(lay, w, D, sp6, so_mem, e_sum, e_last, u, t_delta=schedule.T_DELTA, stream=None):
    """sbr_ilc_update: e_sum, e_last, u [S,n] updated in place from so_mem [S,n] and the set-point memory sp6 [6,n]."""
    lib = _abi.load()
    S, n = so_mem.shape
    pw, _ = core._dev_ptr(w, 1, S, name="w")
    pd, _ = core._dev_ptr(D, 1, S, name="D")
    psp, l0 = core._dev_ptr(sp6, 6, n, name="sp6")
    pso, l1 = core._dev_ptr(so_mem, S, n, name="so_mem")
    pes, l2 = core._dev_ptr(e_sum, S, n, name="e_sum")
    pel, l3 = core._dev_ptr(e_last, S, n, name="e_last")
    pu, l4 = core._dev_ptr(u, S, n, name="u")
    ld = core._same_ld([l0, l1, l2, l3, l4], "ilc_update")
    with torch.cuda.device(so_mem.device):
        rc = lib.sbr_ilc_update(n, ld, C.byref(lay), pw, pd, psp, pso, pes, pel, pu, float(t_delta), KC_B, TAUI_B, TAUD_B,
                                core._stream_ptr(stream))
    _abi.check(rc, "sbr_ilc_update")
    return u


class SbrIlcVecEnv(object):
    """N x the batch-to-batch controlled plant of `SBR-v0`.

    reset(influent=None) -> obs [N,14]: runs cycle 0 (feedback PID only, set-points 2/2/2, from the module's x0) for every
        env, keeps its KLa profile as the feed-forward base and its So / set-point memories for the controller, zeroes
        the controller (gym_SBR_env0.py:58-70,105-135).  obs = (x_last + influent) / scale with obs[0] := 1 (:153-176).
    step(action [N,3]) -> (obs [N,14], reward [N], done [N], info): action = DO set-points of phases 3, 5, 8, clipped to
        [0, 5] (:187); batch-to-batch update, then one feed-forward + feedback cycle from the previous end state with the
        influent drawn after the previous step; done is always True (:206); a new influent is drawn for the next cycle.
    learn: "frozen" (default) = what SbrEnv.step does: its new memories land in local names (:200), so the controller
        learns from cycle 0's So memory for ever; "feedback" = the So / set-point memories of the last cycle are fed back
        (batch-to-batch learning as the module is evidently meant to work; pinned against the same two reference functions).
    influent: [N,14] tensor (row 0 is overwritten with the fill flow 31.4285, :193) or None = per-env draws of
        buffer_tank2.influent.buffer_tank(0, 12) (gym_SBR_env0.py:74,208) from the counter-based device sampler (its mixing
        arithmetic is bit-identical to the reference's for the same normals; rng must be "philox").
    """

    num_actions = 3
    num_obs = 14
    scenario = 0
    influent_tables = "buffer_tank2"

    def __init__(self, num_envs, device="cuda", seed=None, learn="frozen", params=None, rng="philox", env_offset=0,
                 record_feed_forward=False):
        from . import vec_env
        if learn not in ("frozen", "feedback"):
            raise ValueError("learn must be 'frozen' or 'feedback'")
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _abi.SbrLibraryError("SbrIlcVecEnv needs a CUDA device: there is no CPU fallback")
        self.lib = _abi.load()
        self.learn = learn
        self.params = apply_constants(params if params is not None else _abi.default_params())
        self.sched = schedule.cycle_schedule()                 # the reference's output grid: one memory sample per point
        w, D, self.layout = weights(self.sched)
        f = dict(dtype=torch.float64, device=self.device)
        self._w, self._D = torch.as_tensor(w, **f), torch.as_tensor(D, **f)
        if rng != "philox":
            raise ValueError("SbrIlcVecEnv draws its influent from the counter-based sampler only (rng='philox')")
        vec_env._init_rng(self, seed, rng, env_offset)
        n, S = self.num_envs, int(self.layout.n_samples)
        # this path is sized by HBM capacity, not by throughput: refuse early instead of running the device out of memory
        need = (6 + bool(record_feed_forward)) * S * n * 8
        free, _total = torch.cuda.mem_get_info(self.device)
        if need > 0.9 * free:
            raise ValueError("SbrIlcVecEnv(%d envs) needs %.1f GB of sample memories (%d rows of %d samples per env), "
                             "%.1f GB are free on %s" % (n, need / 1e9, 6 + bool(record_feed_forward), S, free / 1e9,
                                                         self.device))
        self.x = torch.tensor(X0_ILC, **f)[:, None].repeat(1, n).contiguous()
        self.influent = torch.zeros((_abi.NX, n), **f)
        self._sp = torch.zeros((3, n), **f)
        self._sp6 = torch.zeros((6, n), **f)
        # sample memories [S][N] (38 kB per env each).  The cycle kernel writes its So memory into `_so_next` while the
        # update kernel reads `so_learn`; learn="feedback" swaps the two after every cycle, so nothing is ever copied.
        # The clamped feed-forward profile the reference returns as Kla_memory (and never reads) is only written on request:
        # forming it costs two more reads and one more write of a sample row per cycle.
        self.kla_base = torch.zeros((S, n), **f)
        self.so_learn = torch.zeros((S, n), **f)               # the So memory the controller reads
        self._so_next = torch.zeros((S, n), **f)
        self.kla_ff = torch.zeros((S, n), **f) if record_feed_forward else None
        self._cyc = IlcCycleOut(n, S, self.device, so_mem=self._so_next, kla_mem=False)
        self.e_sum = torch.zeros((S, n), **f)
        self.e_last = torch.zeros((S, n), **f)
        self.u = torch.zeros((S, n), **f)
        self._obs_scale = torch.tensor(OBS_SCALE, **f)[:, None]
        self._done = torch.ones((n,), dtype=torch.bool, device=self.device)
        self._ready = False

    # -- helpers --
    def _next_influent(self, influent):
        if influent is None:
            from . import vec_env
            vec_env._draw_influent(self, out=self.influent)
        else:
            t = torch.as_tensor(influent, dtype=torch.float64, device=self.device)
            self.influent.copy_(t.t() if t.shape == (self.num_envs, _abi.NX) else t)

    def _obs(self):
        s = (self.x + self.influent) / self._obs_scale          # np.sum of [x_last, influent_mixed] (:153-155)
        s[0] = 1.0
        return s.t().contiguous()

    def reset(self, influent=None):
        n = self.num_envs
        self._next_influent(influent)
        self.influent[0] = FILL_FLOW                             # gym_SBR_env0.py:76
        self.x.copy_(torch.tensor(X0_ILC, dtype=torch.float64, device=self.device)[:, None].expand(-1, n))
        for k in range(3):
            self._sp[k] = CYCLE0_SETPOINTS[k]
        # cycle 0 writes straight into the feed-forward base and the controller's So memory
        self._cyc.so_mem, self._cyc.kla_mem = self.so_learn, self.kla_base
        cycle_ilc(self.x, self.influent, self._sp, self.params, self.sched, self.layout, out=self._cyc)
        self._cyc.so_mem, self._cyc.kla_mem = self._so_next, self.kla_ff
        self.e_sum.zero_(); self.e_last.zero_(); self.u.zero_()
        self.x.copy_(self._cyc.x_last)
        self._ready = True
        return self._obs()

    def step(self, action, influent=None):
        if not self._ready:
            raise RuntimeError("call reset() first")
        a = torch.as_tensor(action, dtype=torch.float64, device=self.device).reshape(self.num_envs, 3)
        a = torch.where(a < ACTION_LOW, torch.full_like(a, ACTION_LOW), torch.where(a > ACTION_HIGH, torch.full_like(a, ACTION_HIGH), a))
        self._sp.copy_(a.t())
        self.influent[0] = FILL_FLOW                             # gym_SBR_env0.py:193
        # set-point memories handed to batch_PID: phases 1, 2, 4 carry their zeros; phases 3, 5, 8 are the previous memory
        # rescaled, sp_prev / sp_prev[0] * action (gym_SBR_env0.py:251-253) -- in the module the previous memory is cycle 0's
        # for ever (set-point 2), so this is the action itself, exactly.  With the last cycle fed back the same expression
        # would turn 0 / 0 into NaN after a zero set-point, a state the reference can never reach; the action is used.
        self._sp6.zero_()
        self._sp6[2], self._sp6[4], self._sp6[5] = self._sp[0], self._sp[1], self._sp[2]
        ilc_update(self.layout, self._w, self._D, self._sp6, self.so_learn, self.e_sum, self.e_last, self.u)
        self._cyc.so_mem, self._cyc.kla_mem = self._so_next, self.kla_ff
        cycle_ilc(self.x, self.influent, self._sp, self.params, self.sched, self.layout, kla_base=self.kla_base, u=self.u,
                  out=self._cyc)
        so_written = self._so_next
        if self.learn == "feedback":
            self.so_learn, self._so_next = self._so_next, self.so_learn
        self.x.copy_(self._cyc.x_last)
        o = self._cyc.out
        info = dict(x_last=self._cyc.x_last, status=self._cyc.status, Qeff=o[_abi.ILC_QEFF], Qw=o[_abi.ILC_QW],
                    OCI=o[_abi.ILC_OCI], kla3_mean=o[_abi.ILC_KLA3_MEAN], kla5_mean=o[_abi.ILC_KLA5_MEAN],
                    kla8_mean=o[_abi.ILC_KLA8_MEAN], so_mem=so_written, kla_ff=self.kla_ff, u_batch=self.u,
                    counters=self._cyc.counters, reward_pinned=False)
        self._next_influent(influent)                            # buffer_tank(0, 12) for the next cycle (:208)
        return self._obs(), o[_abi.ILC_REWARD].clone(), self._done, info

    # -- checkpoint / resume (the reference keeps all of this in module globals and has no resume path) --
    _CKPT = ("x", "influent", "kla_base", "so_learn", "e_sum", "e_last", "u")

    def state_dict(self):
        from . import vec_env
        sd = dict(kind="SBR-v0", num_envs=self.num_envs, learn=self.learn, ready=self._ready,
                  rng_state=vec_env._rng_state(self))
        sd.update({k: getattr(self, k).clone() for k in self._CKPT})
        return sd

    def load_state_dict(self, sd):
        from . import vec_env
        if sd["kind"] != "SBR-v0" or sd["num_envs"] != self.num_envs:
            raise ValueError("checkpoint is for %s with %d envs" % (sd["kind"], sd["num_envs"]))
        for k in self._CKPT:
            getattr(self, k).copy_(sd[k])
        self.learn, self._ready = sd["learn"], bool(sd["ready"])
        vec_env._load_rng_state(self, sd["rng_state"])


class SbrV1VecEnv(object):
    """N x `SBR-v1` (gym_SBR_env1.py:103-203): the plant of `SBR-v0` under its feedback DO-PID alone -- every step is one
    `SBR_model_FBc_implemented.run` (= SBR_model_PID_on.run: each cycle starts from KLa 240, each phase from the previous
    phase's last KLa), from the state the previous step ended in.

    reset(influent=None) -> obs [N,14] of the module's initial state (no cycle is run, :105-126);
    step(action [N,3]) -> (obs, reward, done, info): action = DO set-points of phases 3, 5, 8 clipped to [0, 5] (:131).
    The reference's step() raises on the same seven-argument reward call as `SBR-v0` (:151): the reward here is
    module_reward.sbr_reward's formula on the cycle's applied KLa, by construction; the cycle is pinned against
    SBR_model_FBc_implemented.run (tests/golden/ilc_seed0.npz, `v1_*`).  Influent as for SbrIlcVecEnv."""

    num_actions = 3
    num_obs = 14
    scenario = 0
    influent_tables = "buffer_tank2"

    def __init__(self, num_envs, device="cuda", seed=None, params=None, rng="philox", env_offset=0, order="action"):
        from . import vec_env
        if order not in ("action", "none"):
            raise ValueError("order must be 'action' or 'none'")
        self.order, self._sorted = order, None
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _abi.SbrLibraryError("SbrV1VecEnv needs a CUDA device: there is no CPU fallback")
        self.lib = _abi.load()
        self.params = apply_constants(params if params is not None else _abi.default_params())
        self.sched = schedule.cycle_schedule()
        self.layout = layout(self.sched)
        if rng != "philox":
            raise ValueError("SbrV1VecEnv draws its influent from the counter-based sampler only (rng='philox')")
        vec_env._init_rng(self, seed, rng, env_offset)
        n = self.num_envs
        f = dict(dtype=torch.float64, device=self.device)
        self.x = torch.tensor(X0_ILC, **f)[:, None].repeat(1, n).contiguous()
        self.influent = torch.zeros((_abi.NX, n), **f)
        self._sp = torch.zeros((3, n), **f)
        self._cyc = IlcCycleOut(n, int(self.layout.n_samples), self.device, so_mem=False, kla_mem=False)
        self._obs_scale = torch.tensor(OBS_SCALE, **f)[:, None]
        self._done = torch.ones((n,), dtype=torch.bool, device=self.device)
        self._ready = False

    _next_influent = SbrIlcVecEnv._next_influent
    _obs = SbrIlcVecEnv._obs

    def reset(self, influent=None):
        self._next_influent(influent)
        self.influent[0] = FILL_FLOW
        self.x.copy_(torch.tensor(X0_ILC, dtype=torch.float64, device=self.device)[:, None].expand(-1, self.num_envs))
        self._ready = True
        return self._obs()

    def step(self, action, influent=None):
        if not self._ready:
            raise RuntimeError("call reset() first")
        a = torch.as_tensor(action, dtype=torch.float64, device=self.device).reshape(self.num_envs, 3)
        a = torch.where(a < ACTION_LOW, torch.full_like(a, ACTION_LOW), torch.where(a > ACTION_HIGH, torch.full_like(a, ACTION_HIGH), a))
        self._sp.copy_(a.t())
        self.influent[0] = FILL_FLOW
        c = self._cyc
        if self.order == "action" and self.num_envs > 32:
            # divergence-aware order, as SbrV2VecEnv: this env keeps no per-env memories, so its data can be handed to the
            # adaptive kernel sorted by set-point (one coalesced gather in, one scatter out; results are unaffected)
            if self._sorted is None:
                f = dict(dtype=torch.float64, device=self.device)
                self._sorted = dict(x=torch.empty_like(self.x), influent=torch.empty_like(self.influent),
                                    sp=torch.empty_like(self._sp),
                                    out=IlcCycleOut(self.num_envs, int(self.layout.n_samples), self.device, so_mem=False,
                                                    kla_mem=False))
            z = self._sorted
            a0, a1 = self._sp[0] / ACTION_HIGH, self._sp[1] / ACTION_HIGH
            perm = torch.argsort(torch.floor(a0 * 255.999) + 0.999 * a1)
            core.permute_rows(perm, [(self.x, z["x"]), (self.influent, z["influent"]), (self._sp, z["sp"])])
            zo = cycle_ilc(z["x"], z["influent"], z["sp"], self.params, self.sched, self.layout, out=z["out"])
            core.permute_rows(perm, [(zo.x_last, c.x_last), (zo.out, c.out), (zo.status, c.status),
                                     (zo.counters, c.counters)], scatter=True)
        else:
            cycle_ilc(self.x, self.influent, self._sp, self.params, self.sched, self.layout, out=c)
        self.x.copy_(c.x_last)
        o = c.out
        info = dict(x_last=c.x_last, status=c.status, Qeff=o[_abi.ILC_QEFF], Qw=o[_abi.ILC_QW],
                    OCI=o[_abi.ILC_OCI], kla3_mean=o[_abi.ILC_KLA3_MEAN], kla5_mean=o[_abi.ILC_KLA5_MEAN],
                    kla8_mean=o[_abi.ILC_KLA8_MEAN], counters=c.counters, reward_pinned=False)
        self._next_influent(influent)
        return self._obs(), o[_abi.ILC_REWARD].clone(), self._done, info

    def state_dict(self):
        from . import vec_env
        return dict(kind="SBR-v1", num_envs=self.num_envs, ready=self._ready, x=self.x.clone(),
                    influent=self.influent.clone(), rng_state=vec_env._rng_state(self))

    def load_state_dict(self, sd):
        from . import vec_env
        if sd["kind"] != "SBR-v1" or sd["num_envs"] != self.num_envs:
            raise ValueError("checkpoint is for %s with %d envs" % (sd["kind"], sd["num_envs"]))
        self.x.copy_(sd["x"]); self.influent.copy_(sd["influent"])
        self._ready = bool(sd["ready"])
        vec_env._load_rng_state(self, sd["rng_state"])
