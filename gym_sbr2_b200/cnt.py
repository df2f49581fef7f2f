"""The five remaining interval-per-step ids of the reference on the batched kernels:

  id            reference class   file                          kind
  SBRCnt-v0     SbrCnt0           gym_SBR_continuous0.py        "cnt0"
  SBRCnt-v1     SbrCnt1           gym_SBR_continuous1.py        "cnt1"
  SBRCnt-v2     SbrCnt2           gym_SBR_continuous2.py        "cnt2"
  SBRCntMA-v1   SbrCntMA1         gym_SBR_continuous_MA1.py     "ma1"
  SBROS-v2      SbrOS1            gym_SBR_oneshot1.py           "os2"

`SbrCntVecEnv(kind, N)` steps N reactors per launch (sbr_cnt_reset / sbr_cnt_step, include/sbr_b200.h); the single-env
classes with the reference's names and tuple shapes live in envs/single.py.  In the reference every step() of these
ids raises NameError in module_reward_continuous1.sbr_reward; the reward here is the repaired form disclosed in
oracle/make_golden_cnt.py, everything else follows the unmodified env modules (golden episodes in tests/golden/cnt_*).
"""
import ctypes as C

import torch

from . import _abi, core, schedule
from .vec_env import _draw_influent, _init_rng, _on

KINDS = {"cnt0": _abi.CNT_V0, "cnt1": _abi.CNT_V1, "cnt2": _abi.CNT_V2, "ma1": _abi.CNT_MA1, "os2": _abi.CNT_OS2}
ENV_ID = {"cnt0": "SBRCnt-v0", "cnt1": "SBRCnt-v1", "cnt2": "SBRCnt-v2", "ma1": "SBRCntMA-v1", "os2": "SBROS-v2"}
OBS_ROWS = {"cnt0": 7, "cnt1": 5, "cnt2": 5, "ma1": 5, "os2": 33}
# env.step calls per episode (measured on the reference: fixed schedules)
EPISODE_STEPS = {"cnt0": 466, "cnt1": 228, "cnt2": 228, "ma1": 463, "os2": 463}


def cnt_config(kind):
    """SbrCntConfig with the reference's module-level constants of that file."""
    q = _abi.SbrCntConfig()
    q.kind = KINDS[kind]
    if kind == "cnt0":
        q.Kc_DO, q.tauI_DO, q.tauD_DO = 10.0, 0.5, 0.00005            # gym_SBR_continuous0.py:80-82
    else:
        q.Kc_DO, q.tauI_DO, q.tauD_DO = 100.0, 20.0, 0.0              # gym_SBR_continuous1.py:77-79 (and the others)
    q.Kc_EC, q.tauI_EC, q.tauD_EC = 1.0, 20.0, 0.0                    # gym_SBR_continuous2.py:90-92, gym_SBR_oneshot1.py:91-93
    q.ec_conc = 400000 / 20648.38 * 1.32                              # gym_SBR_continuous2.py:94, gym_SBR_oneshot1.py:95
    q.u_ec_max = 15.0                                                 # gym_SBR_continuous_MA1.py:336-339
    if kind == "ma1":
        q.Kc_EC, q.tauI_EC = 10.0, 0.5                                # gym_SBR_continuous_MA1.py:89-91
        q.ec_conc = 4000 / 20648.38 * 1.32                            # :93
    if kind == "cnt2":
        q.u_ec_max = 5.0                                              # gym_SBR_continuous2.py:329-332
    q.ec_fill_max = 5.0                                               # EC_control_par[5], gym_SBR_continuous2.py:87
    q.u_ec_init = 2.0 if kind in ("cnt2", "ma1", "os2") else 0.0      # gym_SBR_continuous2.py:186
    m = schedule.batch_time_marks()
    q.tm2_0, q.tm2_1, q.tm4_0 = m[1][0], m[1][1], m[3][0]
    return q


class CntBuffers(object):
    def __init__(self, n, device, obs_rows):
        f = dict(dtype=torch.float64, device=device)
        self.n = n
        self.st = torch.zeros((_abi.CNT_ROWS, n), **f)
        self.obs = torch.zeros((obs_rows, n), **f)
        self.reward = torch.zeros((n,), **f)
        self.done = torch.ones((n,), dtype=torch.uint8, device=device)
        self.status = torch.zeros((n,), dtype=torch.int32, device=device)
        self.counters = torch.zeros((2, n), dtype=torch.int32, device=device)


def cnt_reset(cfg, buf, influent, params, sched, x0=None, mask=None, mode=_abi.MODE_DP45, tol=None, stream=None):
    """reset() of the five envs for a batch (sbr_cnt_reset).  influent [14,n] (row 0 = fill flow)."""
    lib = _abi.load()
    n = buf.n
    d = core._dev_ptr
    pst, l0 = d(buf.st, _abi.CNT_ROWS, n, name="st")
    pin, l1 = d(influent, _abi.NX, n, name="influent")
    px0, l2 = d(x0, _abi.NX, n, name="x0")
    pmk, _ = d(mask, 1, n, dtype=torch.uint8, name="mask")
    pob, l3 = d(buf.obs, buf.obs.shape[0], n, name="obs")
    pdn, _ = d(buf.done, 1, n, dtype=torch.uint8, name="done")
    pss, _ = d(buf.status, 1, n, dtype=torch.int32, name="status")
    pct, l4 = d(buf.counters, 2, n, dtype=torch.int32, name="counters")
    ld = core._same_ld([l0, l1, l2 if x0 is not None else None, l3, l4], "cnt_reset")
    tol = tol or _abi.make_tol()
    with torch.cuda.device(buf.st.device):
        rc = lib.sbr_cnt_reset(n, ld, C.byref(cfg), px0, pin, pmk, C.byref(params), C.byref(sched), pst, pob, pdn, pss,
                               pct, int(mode), C.byref(tol), core._stream_ptr(stream))
    _abi.check(rc, "sbr_cnt_reset")
    return buf


def cnt_step(cfg, buf, action, params, sched, mode=_abi.MODE_DP45, tol=None, stream=None):
    """step() of the five envs for a batch (sbr_cnt_step).  action [2,n]: row 1 is read by kind os2 only."""
    lib = _abi.load()
    n = buf.n
    d = core._dev_ptr
    pst, l0 = d(buf.st, _abi.CNT_ROWS, n, name="st")
    pac, l1 = d(action, 2, n, name="action")
    pob, l2 = d(buf.obs, buf.obs.shape[0], n, name="obs")
    prw, _ = d(buf.reward, 1, n, name="reward")
    pdn, _ = d(buf.done, 1, n, dtype=torch.uint8, name="done")
    pss, _ = d(buf.status, 1, n, dtype=torch.int32, name="status")
    pct, l3 = d(buf.counters, 2, n, dtype=torch.int32, name="counters")
    ld = core._same_ld([l0, l1, l2, l3], "cnt_step")
    tol = tol or _abi.make_tol()
    with torch.cuda.device(buf.st.device):
        rc = lib.sbr_cnt_step(n, ld, C.byref(cfg), pst, pac, C.byref(params), C.byref(sched), pob, prw, pdn, pss, pct,
                              int(mode), C.byref(tol), core._stream_ptr(stream))
    _abi.check(rc, "sbr_cnt_step")
    return buf


def cnt_rollout_k(cfg, buf, action, policy, rewards, params, sched, mode=_abi.MODE_DP45, tol=None, stream=None,
                  emit_obs=True, act_log=None, obs_log=None):
    """K = rewards.shape[0] consecutive step() calls in ONE launch with the policy head evaluated in-kernel between them
    (sbr_cnt_rollout_k).  action [2,n] in/out (row 1: kind os2 only); policy: an _abi.SbrPolicyMlp mapping the kind's
    observation rows (7 / 5 / 18) to 1 (os2: 2) actions."""
    lib = _abi.load()
    n = buf.n
    K = rewards.shape[0]
    if rewards.shape != (K, n) or not rewards.is_contiguous():
        raise ValueError("rewards must be contiguous [K, n]")
    d = core._dev_ptr
    pst, l0 = d(buf.st, _abi.CNT_ROWS, n, name="st")
    pac, l1 = d(action, 2, n, name="action")
    prw, l2 = d(rewards, K, n, name="rewards") if K > 1 else (C.c_void_p(rewards.data_ptr()), None)
    pob, l3 = d(buf.obs if emit_obs else None, buf.obs.shape[0], n, name="obs")
    pdn, _ = d(buf.done, 1, n, dtype=torch.uint8, name="done")
    pss, _ = d(buf.status, 1, n, dtype=torch.int32, name="status")
    pct, l4 = d(buf.counters, 2, n, dtype=torch.int32, name="counters")
    pal = pol = None
    l5 = l6 = None
    if act_log is not None:
        if act_log.shape != (K, 2, n) or not act_log.is_contiguous():
            raise ValueError("act_log must be contiguous [K, 2, n]")
        pal, l5 = d(act_log.view(2 * K, n), 2 * K, n, name="act_log")
    if obs_log is not None:
        rows = obs_log.shape[1]
        if obs_log.dim() != 3 or obs_log.shape[0] != K or obs_log.shape[2] != n or not obs_log.is_contiguous():
            raise ValueError("obs_log must be contiguous [K, n_in, n]")
        pol, l6 = d(obs_log.view(rows * K, n), rows * K, n, name="obs_log")
    ld = core._same_ld([l0, l1, l2, l3 if emit_obs else None, l4, l5, l6], "cnt_rollout_k")
    tol = tol or _abi.make_tol()
    with torch.cuda.device(buf.st.device):
        rc = lib.sbr_cnt_rollout_k(n, ld, K, C.byref(cfg), pst, pac, C.byref(policy), C.byref(params), C.byref(sched), pob,
                                   prw, pdn, pss, pct, pal, pol, int(mode), C.byref(tol), core._stream_ptr(stream))
    _abi.check(rc, "sbr_cnt_rollout_k")
    return buf


POLICY_INPUTS = {"cnt0": 7, "cnt1": 5, "cnt2": 5, "ma1": 5, "os2": 18}


class SbrCntVecEnv(object):
    """N x one of `SBRCnt-v0/1/2`, `SBRCntMA-v1`, `SBROS-v2` (kind = "cnt0" | "cnt1" | "cnt2" | "ma1" | "os2").

    reset() -> observation; step(action) -> (observation, reward [N], done [N] bool, info); `os2` follows SBROS-v1:
    reset() -> (obs_DO [N,9], obs_EC [N,9]), step(action [N,2]) -> ((obs_DO, obs_EC), state [N,15], reward, done, info).
    Observations (reference file:line in csrc/sbr_cnt.cuh):
      cnt0  [N,7]  [t, Si, Xbh, Xba, So, Sno, Snh] / [0.5, 30, 2599, 168, 2, 13, 0.005]
      cnt1, cnt2, ma1  [N,5]  [t/0.5, So/8, Snh/30, clip(dSo/8), clip(dSnh/20)]
    Actions: cnt0 / cnt1 / cnt2 / ma1 take [N] or [N,1] = a CHANGE of the set-point in force (the reference clips the
    accumulated set-point, not the action); os2 takes absolute [DO set-point, NO3 set-point].
    Influent: buffer_tank(0) per env and reset (every one of the five files draws scenario 0).
    """

    scenario = 0

    def __init__(self, kind, num_envs, device="cuda", seed=None, mode="dp45", rtol=1e-8, atol=1e-10, max_steps=200,
                 params=None, rng="philox", rk4_sub_interval=0, env_offset=0):
        if kind not in KINDS:
            raise ValueError("kind must be one of %s" % sorted(KINDS))
        self.kind = kind
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _abi.SbrLibraryError("SbrCntVecEnv needs a CUDA device: there is no CPU fallback")
        self.lib = _abi.load()
        self.params = params if params is not None else _abi.default_params()
        self.sched = schedule.os_schedule(rk4_sub_interval=rk4_sub_interval)
        self.cfg = cnt_config(kind)
        self.mode = {"rk4": _abi.MODE_RK4, "dp45": _abi.MODE_DP45}[mode] if isinstance(mode, str) else int(mode)
        self.tol = _abi.make_tol(rtol, atol, max_steps)
        self.max_episode_steps = EPISODE_STEPS[kind]
        self.num_actions = 2 if kind == "os2" else 1
        self._init_rng(seed, rng, env_offset)
        n = self.num_envs
        f = dict(dtype=torch.float64, device=self.device)
        self.buf = CntBuffers(n, self.device, OBS_ROWS[kind])
        self.influent = torch.zeros((_abi.NX, n), **f)
        self._loading = torch.zeros((_abi.NX, n), **f)
        self._action = torch.zeros((2, n), **f)
        self.fill_flow = schedule.os_fill_flow(self.params.Qin)          # influent_mixed[0] = Qin / t_memory1[-1]

    _init_rng, _draw_influent = _init_rng, _draw_influent

    def _obs(self):
        o = self.buf.obs
        if self.kind == "os2":
            return o[0:9].t(), o[9:18].t()
        return o.t()

    def reset(self, influent=None, x0=None, mask=None):
        if mask is not None:
            mask = mask.to(self.device).to(torch.uint8).contiguous()
        if influent is None:
            self._draw_influent(mask=mask, out=self.influent)
        else:
            influent = influent.to(self.device, torch.float64)
            if mask is None:
                self.influent.copy_(influent)
            else:
                self.influent.copy_(torch.where(mask.bool()[None, :], influent, self.influent))
        self._loading.copy_(self.influent)
        self._loading[0] = self.fill_flow
        if x0 is not None:
            x0 = x0.to(self.device, torch.float64).contiguous()
        cnt_reset(self.cfg, self.buf, self._loading, self.params, self.sched, x0=x0, mask=mask, mode=self.mode,
                  tol=self.tol)
        return self._obs()

    def step_async(self, action, stream=None):
        n = self.num_envs
        with _on(stream):
            a = action.to(self.device, torch.float64)
            if self.kind == "os2":
                if a.shape != (n, 2):
                    raise ValueError("action must be [N,2], got %s" % (tuple(a.shape),))
                self._action.copy_(a.t())
            else:
                a = a.reshape(-1)
                if a.shape != (n,):
                    raise ValueError("action must be [N] or [N,1], got %s" % (tuple(action.shape),))
                self._action[0].copy_(a)
            return cnt_step(self.cfg, self.buf, self._action, self.params, self.sched, mode=self.mode, tol=self.tol)

    def step(self, action):
        b = self.buf
        self.step_async(action)
        info = dict(status=b.status, counters=b.counters, t=b.st[_abi.CNT_T], u_do=b.st[_abi.CNT_U_DO],
                    u_ec=b.st[_abi.CNT_U_EC], kla=b.st[_abi.CNT_KLA_LAST], ec=b.st[_abi.CNT_EC_LAST],
                    Qw=b.st[_abi.CNT_QW], episode_return=b.st[_abi.CNT_RETURN], episode_steps=b.st[_abi.CNT_STEPS])
        if self.kind == "os2":
            return self._obs(), b.obs[18:33].t(), b.reward, b.done.bool(), info
        return self._obs(), b.reward, b.done.bool(), info

    def render(self, mode="human", close=False):
        print("Reward for this step: {}".format(self.buf.reward))

    def state_dict(self):
        b = self.buf
        return dict(kind=ENV_ID[self.kind], num_envs=self.num_envs, influent=self.influent.clone(),
                    loading=self._loading.clone(), epoch=self.epoch.clone(), seed=self.seed, env_offset=self.env_offset,
                    buf={k: getattr(b, k).clone() for k in ("st", "obs", "reward", "done", "status", "counters")})

    def load_state_dict(self, sd):
        if sd["kind"] != ENV_ID[self.kind] or sd["num_envs"] != self.num_envs:
            raise ValueError("checkpoint is for %s with %d envs" % (sd["kind"], sd["num_envs"]))
        for k, v in sd["buf"].items():
            getattr(self.buf, k).copy_(v)
        self.influent.copy_(sd["influent"]); self._loading.copy_(sd["loading"]); self.epoch.copy_(sd["epoch"])
        self.seed, self.env_offset = int(sd["seed"]), int(sd["env_offset"])
