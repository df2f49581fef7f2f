"""Vectorised SBR environments: N independent reactors stepped by one CUDA launch.

Buffer ownership: the tensors a `step` / `reset` returns (observations, state, reward, done, and the entries of
`info`) are VIEWS of the env's own device buffers -- the kernels write straight into them and the next `step`
overwrites them (zero copies on the step path).  A rollout that keeps them across steps must `.clone()` or copy
them into its own storage, as `rollout.collect_episode(store=True)` does.

`SbrV2VecEnv` is the batched drop-in for the reference's `SbrEnv2` (id `SBR-v2`, gym_SBR_env2.py:58-193):
one `step` = one whole 12-h cycle (fill, 4 react phases, settle, draw, idle) with the DO->KLa PID inside.
Observation / action / reward semantics are the reference's; tensors are torch CUDA float64.  The reference
keeps its state in module globals, so two reference envs in one process share state;
the semantics here are "N fresh processes each running one env".
"""
import contextlib
import os

import numpy as np
import torch

from . import _abi, core, influent as influent_mod, schedule

X0_INIT = (0.6161484733495801, 30, 0.571098000538576, 1440.01157895393, 31.254221999137, 2599.2714348941,
           168.915006750837, 551.901552960823, 2.16607843793004, 13.3791460027604, 0.00562880208518134,
           0.35996687629947, 1.86916737961228, 3.790463057094611)        # gym_SBR_env2.py:78-80
WV = 1.32                                                                  # gym_SBR_env2.py:33
IV = 0.6161484733495801                                                    # gym_SBR_env2.py:85


def _on(stream):
    """Context that makes `stream` torch's current stream (so that the staging copies are ordered with the kernel
    launched on it); a no-op for stream=None."""
    return torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()


def _init_rng(self, seed, rng, env_offset):
    """Influent randomness.  rng="philox" (default): counter-based draws inside sbr_influent_sample keyed by
    (seed, GLOBAL env index = env_offset + i, episode number of that env) -- an env's influent sequence does not
    depend on the batch size, the rank that owns it or the order of resets, so a sharded run reproduces the
    single-GPU run env for env.  rng="torch": one torch Philox stream per env object (draws depend
    on the batch).  rng="numpy": the reference's own numpy stream, consumed like N sequential reference resets."""
    if rng not in ("philox", "torch", "numpy"):
        raise ValueError("rng must be 'philox', 'torch' or 'numpy'")
    self.rng = rng
    self.env_offset = int(env_offset)
    self.seed = int(seed) if seed is not None else int.from_bytes(os.urandom(8), "little")
    self.epoch = torch.zeros((self.num_envs,), dtype=torch.int64, device=self.device)   # episodes started per env
    self._gen = torch.Generator(device=self.device)
    if seed is not None:
        self._gen.manual_seed(int(seed))
    self._np_rng = np.random.RandomState(seed) if seed is not None else np.random


def _draw_influent(self, mask=None, out=None):
    """Per-env buffer_tank(scenario) draws -> influent_mixed [14,N].  mask (philox only): draw for these envs only,
    the other columns of `out` are left as they are."""
    n = self.num_envs
    scn = getattr(self, "_scenario_arg", None)
    scn = self.scenario if scn is None else scn
    if self.rng == "philox":
        return core.influent_sample(n, self.device, self.seed, env_offset=self.env_offset, scenario=scn,
                                    epoch=self.epoch, mask=mask, out=out,
                                    scenario_out=getattr(self, "_scenario_i32", None),
                                    table_set=getattr(self, "influent_tables", "buffer_tank3"))
    if scn < 0:
        raise ValueError("per-env scenario draws need rng='philox'")
    if self.rng == "numpy":
        # N sequential buffer_tank(scenario) calls on one numpy stream, as N reference resets would make
        d = influent_mod.draws_per_reset(scn)
        r = self._np_rng.randn(n, d, influent_mod.N_POINTS)[:, -1, :]
        rnd = torch.as_tensor(np.ascontiguousarray(r.T), dtype=torch.float64).to(self.device)
    else:
        rnd = torch.randn((influent_mod.N_POINTS, n), dtype=torch.float64, device=self.device, generator=self._gen)
    drawn = core.influent_mix(scn, core.soa1(rnd))
    if mask is not None:
        self.epoch += mask.to(torch.int64)
        if out is not None:
            drawn = torch.where(mask.bool()[None, :], drawn, out)
    else:
        self.epoch += 1
    if out is not None:
        out.copy_(drawn)
        return out
    return drawn


class SbrV2VecEnv(object):
    """N x `SBR-v2`.  reset() -> obs [N,3]; step(action [N,3]) -> (obs [N,3], reward [N], done [N], info).

    action: raw, clipped to [0,1] in the kernel; DO set-points of phases 3, 5, 8 = 8*action
            (gym_SBR_env2.py:133,184-186).
    obs after reset: [V, (sum COD - 5145)/10, Snh/30] of the element-wise SUM x0 + influent_mixed
            (gym_SBR_env2.py:108-118); after step: [Qeff, COD_eff, Snh_eff/30] (:162-168).
    done is always True (one step = one episode, :160); like the reference, every step replays the cycle from
    x0_init with the influent drawn at the last reset (:88-99,152-153).
    info: x_last [14,N], status [N], counters [2,N], aux rows by name (OCI, Qw, EQI, eff_*, kla*_mean).
    """

    num_actions = 3
    num_obs = 3
    scenario = 0                     # buffer_tank(0), gym_SBR_env2.py:104

    def __init__(self, num_envs, device="cuda", seed=None, mode="rk4", rtol=1e-8, atol=1e-10, max_steps=200,
                 params=None, rng="philox", substeps=None, order="auto", env_offset=0, action_kind="do_setpoint"):
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _abi.SbrLibraryError("SbrV2VecEnv needs a CUDA device: there is no CPU fallback")
        self.lib = _abi.load()
        self.params = params if params is not None else _abi.default_params()
        self.sched = schedule.cycle_schedule(substeps=substeps)     # RK4 sub-steps per interval (None = reference grid)
        self.mode = {"rk4": _abi.MODE_RK4, "dp45": _abi.MODE_DP45}[mode] if isinstance(mode, str) else int(mode)
        # action_kind="kla": the three actions are the KLa of phases 3, 5, 8 as fractions of 240 1/d, the DO controller
        # is bypassed (BASELINE configs[1] "random KLa actions"; no reference env takes a raw KLa -- see the header)
        if action_kind not in ("do_setpoint", "kla"):
            raise ValueError("action_kind must be 'do_setpoint' or 'kla'")
        self.action_kind = action_kind
        self.tol = _abi.make_tol(rtol, atol, max_steps, flags=_abi.FLAG_RAW_KLA if action_kind == "kla" else 0)
        # divergence-aware ordering: with adaptive steps, envs are handed to the kernel in the order of their first DO
        # set-point (the per-env step count is a function of it), physically reordered by one gather launch before
        # and one scatter launch after the cycle kernel so that its loads and stores stay unit-stride; results are
        # unaffected and the caller's buffers keep env order
        if order not in ("auto", "action", "none"):
            raise ValueError("order must be 'auto', 'action' or 'none'")
        self.order = ("action" if self.mode == _abi.MODE_DP45 else "none") if order == "auto" else order
        self._init_rng(seed, rng, env_offset)
        n = self.num_envs
        f = dict(dtype=torch.float64, device=self.device)
        self.x0 = torch.tensor(X0_INIT, **f)[:, None].repeat(1, n).contiguous()
        self.influent = torch.zeros((_abi.NX, n), **f)
        self.Qin = WV - IV
        t_fill = schedule.T_CYCLE * schedule.T_RATIO[0]
        self.fill_flow = self.Qin / t_fill                                    # gym_SBR_env2.py:144
        self._loading = torch.zeros((_abi.NX, n), **f)
        self._action = torch.zeros((3, n), **f)
        self._out = core.CycleV2Out(n, self.device)
        self._done = torch.ones((n,), dtype=torch.bool, device=self.device)
        self._sorted = None                  # gather targets of the ordered launch, allocated on first use

    _init_rng, _draw_influent = _init_rng, _draw_influent

    def reset(self, influent=None, x0=None):
        """influent: optional [14,N] influent_mixed (row 0 ignored at step time); x0: optional [14,N]."""
        if influent is None:
            self._draw_influent(out=self.influent)
        else:
            self.influent.copy_(influent.to(self.device, torch.float64))
        if x0 is not None:
            self.x0.copy_(x0.to(self.device, torch.float64))
        # what the fill phase loads: the influent concentrations with row 0 = the fill flow (gym_SBR_env2.py:144);
        # staged here, once per reset, not per step
        self._loading.copy_(self.influent)
        self._loading[0] = self.fill_flow
        s = self.x0 + self.influent
        cod = s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7]
        return torch.stack([s[0], (cod - 5145) / 10, s[10] / 30], dim=1)

    # -- step ----------------------------------------------------------------------------------------
    def step_async(self, action, stream=None):
        """Launch one cycle for every env; returns the SoA output buffers without synchronising."""
        if action.shape != (self.num_envs, 3):
            raise ValueError("action must be [N,3], got %s" % (tuple(action.shape),))
        with _on(stream):
            self._action.copy_(action.to(self.device, torch.float64).t())
        return self.step_soa(self._action, stream=stream)

    def step_soa(self, action_soa, stream=None):
        """Zero-copy variant: action_soa is the kernel's own layout [3,N] (float64, CUDA, contiguous)."""
        if action_soa.shape != (3, self.num_envs):
            raise ValueError("action_soa must be [3,N], got %s" % (tuple(action_soa.shape),))
        n = self.num_envs
        if self.order != "action" or n <= 32:
            return core.cycle_v2(self.x0, self._loading, action_soa, self.params, self.sched, out=self._out,
                                 mode=self.mode, tol=self.tol, stream=stream)
        if self._sorted is None:
            f = dict(dtype=torch.float64, device=self.device)
            self._sorted = dict(x0=torch.empty((_abi.NX, n), **f), loading=torch.empty((_abi.NX, n), **f),
                                action=torch.empty((3, n), **f), out=core.CycleV2Out(n, self.device))
        z, o = self._sorted, self._out
        with _on(stream):
            # divergence-aware order: envs binned by the first DO set-point (which decides the step count of the long
            # aerobic phase), ordered by the second one inside a bin -- 30.96 of 32 lanes against 30.59 for the first
            # set-point alone and 19.4 in env order (profiles/r02i_sort_key_candidates.log)
            a0, a1 = action_soa[0].clamp(0.0, 1.0), action_soa[1].clamp(0.0, 1.0)
            perm = torch.argsort(torch.floor(a0 * 255.999) + 0.999 * a1)
            core.permute_rows(perm, [(self.x0, z["x0"]), (self._loading, z["loading"]), (action_soa, z["action"])])
            zo = core.cycle_v2(z["x0"], z["loading"], z["action"], self.params, self.sched, out=z["out"],
                               mode=self.mode, tol=self.tol)
            core.permute_rows(perm, [(zo.x_last, o.x_last), (zo.obs, o.obs), (zo.reward, o.reward), (zo.aux, o.aux),
                                     (zo.status, o.status), (zo.counters, o.counters)], scatter=True)
        return o

    def step(self, action):
        o = self.step_async(action)
        info = dict(x_last=o.x_last, status=o.status, counters=o.counters)
        for k, name in enumerate(_abi.AUX_NAMES):
            info[name] = o.aux[k]
        return o.obs.t(), o.reward, self._done, info

    def trajectory(self, action):
        """The cycle `step(action)` runs, with its trajectory: what SBR_model_FB.run returns as `t`, `x` (SBR_model_FB.py:71-86)
        and as its per-interval KLa arrays, sampled at the END of every PID interval of phases 1-5 and 8 (528 records) plus
        the post-draw state (record 492, KLa 0).  For analysis and plotting, off the timed path (sbr_cycle_v2_traj: 67 kB per
        env; it runs interval by interval, so in adaptive mode its steps are not the three-segment kernel's -- same results
        within the tolerance).  Returns dict(t [R,N], x [R,14,N], kla [R,N], x_last [14,N], reward [N], obs [N,3])."""
        if action.shape != (self.num_envs, 3):
            raise ValueError("action must be [N,3], got %s" % (tuple(action.shape),))
        a = action.to(self.device, torch.float64).t().contiguous()
        t_start = [b[0] for b in schedule.phase_bounds()]
        out, traj = core.cycle_v2_traj(self.x0, self._loading, a, self.params, self.sched, t_start, mode=self.mode,
                                       tol=self.tol)
        return dict(t=traj[:, _abi.TRAJ2_T], x=traj[:, _abi.TRAJ2_X:_abi.TRAJ2_X + _abi.NX], kla=traj[:, _abi.TRAJ2_KLA],
                    x_last=out.x_last, reward=out.reward, obs=out.obs.t(), status=out.status)

    def render(self, mode="human", close=False):
        print("Reward for this episode: {}".format(self._out.reward))


class SbrOsVecEnv(object):
    """N x `SBROS-v1` (the reference's `SbrOS`, gym_SBR_oneshot.py:98-2644): one `step` = one 72-s PID interval
    (two when the step crosses a phase boundary); 463 steps per episode; the last one also settles, draws and
    idles.  One kernel launch per `step` for the whole batch; the per-env controller histories the reference keeps
    in module-level lists live in a [35, N] SoA state tensor on the device.

    reset()  -> (obs_DO [N,9], obs_EC [N,9])
    step(a)  -> ((obs_DO [N,9], obs_EC [N,9]), state [N,15], reward [N], done [N] bool, info)   (the reference's
                5-tuple, gym_SBR_oneshot.py:1273)
    action [N,2]: a[:,0] = DO set-point (g/m3, clipped to [0,8], used in aerobic phases), a[:,1] = NO3 set-point
                (clipped to [0,15], used in anoxic phases) (:862-906).  The declared action_space Box([-1],[1]) of
                the reference does not describe what `step` consumes; neither is enforced.
    autoreset: envs whose episode has ended are restarted (new influent draw, reset kernel) at the start of the next
               `step` and then take that step like every other env.  Convention: info["restarted"] marks them and
               info["reset_obs"] holds the observations the reset produced (copies, valid for the marked envs; None
               when nothing restarted) -- the action an on-policy caller passed for a marked env was computed from
               the previous episode's terminal observation, so the first transition of the new episode should be
               formed from reset_obs (or dropped).  While all envs stem from one full reset they end together and
               nothing is drawn or launched in between (one `done.all()` per episode); after a masked reset the
               general path runs a masked influent draw and a masked reset kernel before every step.
               Without autoreset, stepping a finished env is a no-op (reward 0, status SBR_ST_DONE).
    """

    num_actions = 2
    scenario = 6                     # buffer_tank(6), gym_SBR_oneshot.py:180
    max_episode_steps = 463

    def __init__(self, num_envs, device="cuda", seed=None, mode="dp45", rtol=1e-8, atol=1e-10, max_steps=200,
                 params=None, rng="philox", autoreset=False, rk4_sub_interval=0, env_offset=0,
                 emit=("obs_do", "obs_ec", "state"), record_trajectory=False):
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _abi.SbrLibraryError("SbrOsVecEnv needs a CUDA device: there is no CPU fallback")
        # which observation outputs the step kernel writes: a rollout whose policy reads only obs_DO / obs_EC passes
        # emit=("obs_do", "obs_ec") and saves the 120 B per env-step of the 15-dim `state` (it then keeps its reset value)
        self.emit = tuple(emit)
        self.lib = _abi.load()
        self.params = params if params is not None else _abi.default_params()
        self.sched = schedule.os_schedule(rk4_sub_interval=rk4_sub_interval)
        self.mode = {"rk4": _abi.MODE_RK4, "dp45": _abi.MODE_DP45}[mode] if isinstance(mode, str) else int(mode)
        self.tol = _abi.make_tol(rtol, atol, max_steps)
        self.autoreset = bool(autoreset)
        self._init_rng(seed, rng, env_offset)
        n = self.num_envs
        f = dict(dtype=torch.float64, device=self.device)
        self.buf = core.OsBuffers(n, self.device)
        self.influent = torch.zeros((_abi.NX, n), **f)
        self._loading = torch.zeros((_abi.NX, n), **f)
        self._action = torch.zeros((2, n), **f)
        self.fill_flow = schedule.os_fill_flow(self.params.Qin)               # gym_SBR_oneshot.py:287
        # optional trajectory dump (the reference's trajectory(), gym_SBR_oneshot.py:1275-1288): one record per PID
        # interval, written by the step kernel (sbr_os_step_traj).  184 B per env and interval: for analysis and
        # plotting of small batches, not for the timed path.
        self.traj_cap = (470 if record_trajectory is True else int(record_trajectory)) if record_trajectory else 0
        self.traj = torch.full((self.traj_cap, _abi.TRAJ_ROWS, n), float("nan"), **f) if self.traj_cap else None
        self.traj0 = torch.zeros((1 + _abi.NX, n), **f) if self.traj_cap else None      # [t, x] after the fill
        self._init_lockstep()

    _init_rng, _draw_influent = _init_rng, _draw_influent

    def trajectory(self, env=0):
        """Trajectory of env `env` since its last reset, sampled at the ends of the PID intervals (needs
        record_trajectory=...).  dict of numpy arrays: t [M+1], x [M+1,14] (entry 0 = state after the fill phase),
        and per interval kla, ec, u_do, u_ec [M]; per env.step reward, reward_EQI, reward_OCI, reward_AE, reward_EC and
        step_end (index into t / x of the step's last interval)."""
        if self.traj is None:
            raise RuntimeError("construct the env with record_trajectory=True (or a record capacity)")
        T = _abi
        rec = self.traj[:, :, env].cpu().numpy()
        m = int(np.isfinite(rec[:, T.TRAJ_T]).sum())
        rec = rec[:m]
        first = self.traj0[:, env].cpu().numpy()
        has_r = np.isfinite(rec[:, T.TRAJ_REWARD])
        return dict(t=np.concatenate([first[:1], rec[:, T.TRAJ_T]]),
                    x=np.concatenate([first[None, 1:], rec[:, T.TRAJ_X:T.TRAJ_X + T.NX]], axis=0),
                    kla=rec[:, T.TRAJ_KLA], ec=rec[:, T.TRAJ_EC], u_do=rec[:, T.TRAJ_U_DO], u_ec=rec[:, T.TRAJ_U_EC],
                    reward=rec[has_r, T.TRAJ_REWARD], reward_EQI=rec[has_r, T.TRAJ_EQI], reward_OCI=rec[has_r, T.TRAJ_OCI],
                    reward_AE=rec[has_r, T.TRAJ_AE], reward_EC=rec[has_r, T.TRAJ_ECO],
                    step_end=np.nonzero(has_r)[0] + 1)

    def reset(self, influent=None, x0=None, mask=None):
        """influent: optional [14,N] influent_mixed (row 0 is replaced by the fill flow); x0: optional [14,N];
        mask: optional [N] bool/uint8 -- restart only these envs."""
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
        self._note_reset(mask)
        self._loading.copy_(self.influent)
        self._loading[0] = self.fill_flow
        if x0 is not None:
            x0 = x0.to(self.device, torch.float64).contiguous()
        core.os_reset(self.buf, self._loading, self.params, self.sched, x0=x0, mask=mask, mode=self.mode,
                      tol=self.tol)
        if self.traj is not None:
            first = torch.cat([self.buf.st[_abi.OS_T:_abi.OS_T + 1], self.buf.st[:_abi.NX]], dim=0)
            if mask is None:
                self.traj.fill_(float("nan"))
                self.traj0.copy_(first)
            else:
                m = mask.bool()
                self.traj[:, :, m] = float("nan")
                self.traj0.copy_(torch.where(m[None, :], first, self.traj0))
        return self.buf.obs_do.t(), self.buf.obs_ec.t()

    def step_async(self, action, stream=None):
        if action.shape != (self.num_envs, 2):
            raise ValueError("action must be [N,2], got %s" % (tuple(action.shape),))
        with _on(stream):
            self._action.copy_(action.to(self.device, torch.float64).t())
        self._host_steps += 1
        return core.os_step(self.buf, self._action, self.params, self.sched, mode=self.mode, tol=self.tol,
                            stream=stream, emit=self.emit, traj=self.traj)

    def step_k(self, actions_soa, rewards, stream=None):
        """K consecutive env.steps in ONE launch (sbr_os_step_k): actions_soa [K,2,N], rewards [K,N] (out, row k = the
        reward of step k; 0 after the episode has ended).  The observation buffers hold the observation after the last
        step that ran.  Frame-skip / open-loop set-point sequences: the state never leaves the registers in between."""
        if actions_soa.dim() != 3 or actions_soa.shape[1:] != (2, self.num_envs):
            raise ValueError("actions_soa must be [K,2,N], got %s" % (tuple(actions_soa.shape),))
        self._host_steps += int(actions_soa.shape[0])
        self._lockstep = self._lockstep and not self.autoreset
        return core.os_step(self.buf, actions_soa, self.params, self.sched, mode=self.mode, tol=self.tol,
                            stream=stream, emit=self.emit, rewards=rewards, traj=self.traj)

    def step_soa(self, action_soa, stream=None):
        """Zero-copy variant for device-side policies: action_soa is the kernel's own layout [2,N] (float64, CUDA,
        contiguous) and the observations are read from self.buf.obs_do / obs_ec / state ([9,N], [9,N], [15,N]),
        reward from self.buf.reward, done from self.buf.done -- no transposes on either side of the launch."""
        if action_soa.shape != (2, self.num_envs):
            raise ValueError("action_soa must be [2,N], got %s" % (tuple(action_soa.shape),))
        self._host_steps += 1
        return core.os_step(self.buf, action_soa, self.params, self.sched, mode=self.mode, tol=self.tol,
                            stream=stream, emit=self.emit, traj=self.traj)

    def step(self, action):
        b = self.buf
        restarted = reset_obs = None
        if self.autoreset:
            restarted = self._autoreset()
            if restarted is not self._no_restart:
                reset_obs = (b.obs_do.t().clone(), b.obs_ec.t().clone())
        self.step_async(action)
        info = dict(status=b.status, counters=b.counters, t=b.st[_abi.OS_T], Qw=b.st[_abi.OS_QW],
                    episode_return=b.st[_abi.OS_RETURN], episode_steps=b.st[_abi.OS_STEPS], restarted=restarted,
                    reset_obs=reset_obs)
        return (b.obs_do.t(), b.obs_ec.t()), b.state.t(), b.reward, b.done.bool(), info

    def render(self, mode="human", close=False):
        print("Reward for this step: {}".format(self.buf.reward))


class SbrV4VecEnv(object):
    """N x `SBR-v4` (the reference's `SbrEnv4`, gym_SBR_env4.py:71-1294): one `step` = one 72-s PID interval, the
    FILL phase included (26 fill + 466 react steps, then one step that settles, draws and idles: 493 per episode).
    action [N] or [N,1]: change of the DO set-point, accumulated into u and clipped to [0, 8] (:209-218).
    reset() -> state [N,14] (flow-weighted mix of influent and reactor / x_1); step(a) -> (state [N,14] = x / x_1,
    reward [N], done [N] bool, info) -- the standard Gym 4-tuple.  Every reset draws a scenario uniformly from the 8
    influent scenarios per env (np.random.choice(8, 1), :104).

    The reference's step() raises TypeError on numpy >= 1.18 (float `num` in np.linspace); parity is against the
    unmodified source run with numpy < 1.18 linspace semantics (oracle/make_golden_v4.py)."""

    num_actions = 1
    max_episode_steps = 493

    def __init__(self, num_envs, device="cuda", seed=None, mode="dp45", rtol=1e-8, atol=1e-10, max_steps=200,
                 params=None, autoreset=False, rk4_sub_interval=0, env_offset=0, order="auto"):
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _abi.SbrLibraryError("SbrV4VecEnv needs a CUDA device: there is no CPU fallback")
        self.lib = _abi.load()
        self.params = params if params is not None else _abi.default_params()
        self.sched = schedule.os_schedule(rk4_sub_interval=rk4_sub_interval)
        self.mode = {"rk4": _abi.MODE_RK4, "dp45": _abi.MODE_DP45}[mode] if isinstance(mode, str) else int(mode)
        self.tol = _abi.make_tol(rtol, atol, max_steps)
        self.autoreset = bool(autoreset)
        self._init_rng(seed, "philox", env_offset)
        n = self.num_envs
        f = dict(dtype=torch.float64, device=self.device)
        self.buf = core.V4Buffers(n, self.device)
        self.influent = torch.zeros((_abi.NX, n), **f)          # influent_mixed with row 0 = 0.66
        self._scenario_i32 = torch.zeros((n,), dtype=torch.int32, device=self.device)   # written by the sampler
        self._loading = torch.zeros((_abi.NX, n), **f)          # ... with row 0 = the fill flow
        self._action = torch.zeros((n,), **f)
        self.fill_flow = schedule.os_fill_flow(self.params.Qin)               # gym_SBR_env4.py:193
        # divergence-aware placement (adaptive mode): the envs of this path need 1 to 25 integrator steps per interval,
        # persistently per env (correlation 0.9 from one step to the next), and a warp pays for its slowest env.  The
        # persistent state is therefore kept in SLOT order -- slots re-sorted by the previous step's RHS count every few
        # steps (one argsort + one sbr_permute_rows of the state) -- while actions, observations, rewards ... stay
        # indexed by env: the kernels take the slot -> env map (`order`).  Results are bit-identical either way.
        if order not in ("auto", "steps", "none"):
            raise ValueError("order must be 'auto', 'steps' or 'none'")
        self.order = ("steps" if self.mode == _abi.MODE_DP45 and n > 64 else "none") if order == "auto" else order
        self._slot_env = None                   # int32 [N] slot -> env, None = identity
        self._st_alt = None                     # gather target of a re-sort (swapped with buf.st)
        self._init_lockstep()

    def _resort_due(self):
        k = self._host_steps
        return self.order == "steps" and ((k < 64 and k % 8 == 7) or (k >= 64 and k % 32 == 31))

    # envs are placed in groups of `place_group` consecutive envs (one 32-byte sector of every env-indexed row): a slot
    # order that scatters single envs turns each of the kernel's env-indexed loads and stores into a partial-sector
    # access (measured: 0.32 -> 0.64 ms per step), whole sectors cost what unit-stride rows cost
    place_group = 4

    def _resort(self):
        """Re-sort the state slots, group-wise, by the RHS count of the step that just ran."""
        b = self.buf
        G, n = self.place_group, self.num_envs
        ng = n // G
        cnt = b.counters[0]
        key = cnt if self._slot_env is None else cnt[self._slot_env.long()]
        gkey = key[: ng * G].view(ng, G).max(dim=1).values
        gperm = torch.argsort(gkey)
        perm = (gperm[:, None] * G + torch.arange(G, device=self.device)[None, :]).reshape(-1)
        if ng * G < n:
            perm = torch.cat([perm, torch.arange(ng * G, n, device=self.device)])
        if self._st_alt is None:
            self._st_alt = torch.empty_like(b.st)
        core.permute_rows(perm, [(b.st, self._st_alt)])
        b.st, self._st_alt = self._st_alt, b.st
        self._slot_env = perm.to(torch.int32) if self._slot_env is None else self._slot_env[perm]

    def unsort(self):
        """Put the state back into env order (checkpoints, direct access to buf.st)."""
        if self._slot_env is not None:
            b = self.buf
            if self._st_alt is None:
                self._st_alt = torch.empty_like(b.st)
            core.permute_rows(self._slot_env.long(), [(b.st, self._st_alt)], scatter=True)
            b.st, self._st_alt = self._st_alt, b.st
            self._slot_env = None

    def _st_row(self, row):
        """Row `row` of the persistent state in env order."""
        r = self.buf.st[row]
        if self._slot_env is None:
            return r
        out = torch.empty_like(r)
        out[self._slot_env.long()] = r
        return out

    _scenario_arg = -1               # np.random.choice(8, 1) per reset, gym_SBR_env4.py:104: drawn by the sampler
    _init_rng, _draw_influent = _init_rng, _draw_influent

    @property
    def scenario(self):
        """[N] influent scenario (0..7) of every env's current episode."""
        return self._scenario_i32

    def reset(self, influent=None, x0=None, mask=None, scenario=None):
        """influent: optional [14,N] (then `scenario`, optional [N], only labels it); default: per env a scenario
        ~ U{0..7} and one buffer_tank(scenario) draw, both inside one sbr_influent_sample launch (masked envs only)."""
        if mask is not None:
            mask = mask.to(self.device).to(torch.uint8).contiguous()
        if influent is None:
            self._draw_influent(mask=mask, out=self.influent)
        else:
            influent = influent.to(self.device, torch.float64)
            scn = None if scenario is None else scenario.to(self.device, torch.int32)
            if mask is None:
                self.influent.copy_(influent)
                if scn is not None:
                    self._scenario_i32.copy_(scn)
            else:
                self.influent.copy_(torch.where(mask.bool()[None, :], influent, self.influent))
                if scn is not None:
                    self._scenario_i32.copy_(torch.where(mask.bool(), scn, self._scenario_i32))
        self._note_reset(mask)
        self._loading.copy_(self.influent)
        self._loading[0] = self.fill_flow
        if x0 is not None:
            x0 = x0.to(self.device, torch.float64).contiguous()
        if mask is None:
            self._slot_env = None                # a full reset rewrites every slot: back to the identity placement
        core.v4_reset(self.buf, self._loading, self.params, x0=x0, mask=mask, order=self._slot_env)
        return self.buf.obs.t()

    def step_async(self, action, stream=None):
        action = action.reshape(-1)
        if action.shape != (self.num_envs,):
            raise ValueError("action must be [N] or [N,1], got %s" % (tuple(action.shape),))
        with _on(stream):
            self._action.copy_(action.to(self.device, torch.float64))
            out = core.v4_step(self.buf, self._loading, self._action, self.params, self.sched, mode=self.mode,
                               tol=self.tol, order=self._slot_env)
            if self._resort_due():
                self._resort()
        self._host_steps += 1
        return out

    def step(self, action):
        b = self.buf
        restarted = reset_obs = None
        if self.autoreset:
            restarted = self._autoreset()
            if restarted is not self._no_restart:
                reset_obs = b.obs.t().clone()
        self.step_async(action)
        info = dict(status=b.status, counters=b.counters, t=self._st_row(_abi.V4_T), u=self._st_row(_abi.V4_U),
                    Qw=self._st_row(_abi.V4_QW), episode_return=self._st_row(_abi.V4_RETURN),
                    episode_steps=self._st_row(_abi.V4_STEPS), restarted=restarted, reset_obs=reset_obs)
        return b.obs.t(), b.reward, b.done.bool(), info

    def render(self, mode="human", close=False):
        print("Reward for this step: {}".format(self.buf.reward))


# ---------------------------------------------------------------------------------------------------------
# autoreset without per-step work.  Episodes have a fixed length (463 / 493 steps), so as long as every env was
# started by the same full reset the host knows when they end: nothing is drawn or launched on the other steps and
# the end is confirmed with one `done.all()` per episode.  A masked (partial) reset, a checkpoint load or a failed
# confirmation switch to the general path: masked reset kernel + fresh influent draw before every step.
# ---------------------------------------------------------------------------------------------------------
def _init_lockstep(self):
    self._lockstep = False
    self._host_steps = 0
    self._no_restart = torch.zeros((self.num_envs,), dtype=torch.uint8, device=self.device)


def _note_reset(self, mask):
    if mask is None:
        self._lockstep, self._host_steps = True, 0
    else:
        self._lockstep = False


def _autoreset(self):
    b = self.buf
    if self._lockstep:
        if self._host_steps < self.max_episode_steps:
            return self._no_restart
        if bool(b.done.all()):                  # the one synchronisation per episode
            restarted = b.done.clone()
            self.reset()
            return restarted
        self._lockstep = False
    # general path: masked influent draw + masked reset kernel before every step (both return at once for running
    # envs).  Every restart draws afresh: the sampler's counter is (global env index, that env's episode number).
    restarted = b.done.clone()
    self.reset(mask=restarted)
    return restarted


for _cls in (SbrOsVecEnv, SbrV4VecEnv):
    _cls._init_lockstep, _cls._note_reset, _cls._autoreset = _init_lockstep, _note_reset, _autoreset


# ---------------------------------------------------------------------------------------------------------
# checkpoint / resume: the whole simulator state is a handful of torch tensors (the reference keeps it in module
# globals and has no resume path); `torch.save(env.state_dict(), path)` is the checkpoint.
# ---------------------------------------------------------------------------------------------------------
def _rng_state(self):
    np_state = self._np_rng.get_state() if isinstance(self._np_rng, np.random.RandomState) else None
    return dict(rng=self.rng, seed=self.seed, env_offset=self.env_offset, epoch=self.epoch.clone(),
                gen=self._gen.get_state(), np_state=np_state)


def _load_rng_state(self, sd):
    self.rng, self.seed, self.env_offset = sd["rng"], int(sd["seed"]), int(sd["env_offset"])
    self.epoch.copy_(sd["epoch"])
    self._gen.set_state(sd["gen"].cpu())
    if sd.get("np_state") is not None:
        if not isinstance(self._np_rng, np.random.RandomState):
            self._np_rng = np.random.RandomState()
        self._np_rng.set_state(sd["np_state"])


def _state_dict_v2(self):
    return dict(kind="SBR-v2", num_envs=self.num_envs, x0=self.x0.clone(), influent=self.influent.clone(),
                loading=self._loading.clone(), rng_state=_rng_state(self))


def _load_state_dict_v2(self, sd):
    if sd["kind"] != "SBR-v2" or sd["num_envs"] != self.num_envs:
        raise ValueError("checkpoint is for %s with %d envs" % (sd["kind"], sd["num_envs"]))
    self.x0.copy_(sd["x0"]); self.influent.copy_(sd["influent"]); self._loading.copy_(sd["loading"])
    _load_rng_state(self, sd["rng_state"])


# every device buffer a policy-driven rollout reads before the next step writes it (observations, reward, done,
# status, counters) is part of the checkpoint, not only the integrator state
_BUF_FIELDS = ("st", "done", "obs_do", "obs_ec", "state", "obs", "reward", "status", "counters")


def _state_dict_buf(kind):
    def state_dict(self):
        if hasattr(self, "unsort"):
            self.unsort()                        # checkpoints hold the state in env order
        b = self.buf
        sd = dict(kind=kind, num_envs=self.num_envs, influent=self.influent.clone(), loading=self._loading.clone(),
                  rng_state=_rng_state(self), lockstep=self._lockstep, host_steps=self._host_steps,
                  buf={k: getattr(b, k).clone() for k in _BUF_FIELDS if hasattr(b, k)})
        if hasattr(self, "_scenario_i32"):
            sd["scenario"] = self._scenario_i32.clone()
        return sd

    def load_state_dict(self, sd):
        if sd["kind"] != kind or sd["num_envs"] != self.num_envs:
            raise ValueError("checkpoint is for %s with %d envs" % (sd["kind"], sd["num_envs"]))
        if hasattr(self, "unsort"):
            self._slot_env = None
        for k, v in sd["buf"].items():
            getattr(self.buf, k).copy_(v)
        self._lockstep, self._host_steps = bool(sd["lockstep"]), int(sd["host_steps"])
        self.influent.copy_(sd["influent"]); self._loading.copy_(sd["loading"])
        _load_rng_state(self, sd["rng_state"])
        if sd.get("scenario") is not None:
            self._scenario_i32.copy_(sd["scenario"])
    return state_dict, load_state_dict


SbrV2VecEnv.state_dict, SbrV2VecEnv.load_state_dict = _state_dict_v2, _load_state_dict_v2
SbrOsVecEnv.state_dict, SbrOsVecEnv.load_state_dict = _state_dict_buf("SBROS-v1")
SbrV4VecEnv.state_dict, SbrV4VecEnv.load_state_dict = _state_dict_buf("SBR-v4")
