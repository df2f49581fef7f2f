"""Per-instance Gym envs = batch-of-one views of the vector envs (drop-in for the reference's odeint path)."""
import numpy as np
import torch

from .. import influent as influent_mod
from ..spaces import Box, env_base
from ..cnt import SbrCntVecEnv
from ..ilc import ACTION_HIGH, ACTION_LOW, SbrIlcVecEnv, SbrV1VecEnv
from ..vec_env import SbrOsVecEnv, SbrV2VecEnv, SbrV4VecEnv

_Base = env_base()


def _device(device):
    if device is None:
        device = "cuda:0"
    return torch.device(device)


class SbrEnv2(_Base):
    """`SBR-v2` (gym_SBR_env2.py:58-193): one step = one whole 12-h cycle; action = 3 values in [0,1] -> DO
    set-points 8*a of phases 3, 5, 8; obs = [Qeff, COD_eff, Snh_eff/30]; done = True after every step."""
    metadata = {"render.modes": ["human"]}

    def __init__(self, device=None, mode="rk4", rtol=1e-8, atol=1e-10):
        self.action_space = Box(np.array([0., 0., 0.]), np.array([1.0, 1.0, 1.0]), dtype=np.float32)       # :64
        self.observation_space = Box(low=np.array([0.5, 0, 0]), high=np.array([1.33, 2.5, 2]), dtype=np.float32)
        self.reward = 0
        self._vec = SbrV2VecEnv(1, device=_device(device), mode=mode, rtol=rtol, atol=atol)
        self.influent_mixed = None
        self.info = {}

    def reset(self):
        # one buffer_tank(0) call on the GLOBAL numpy RNG, bit-exact with the reference (buffer_tank3.py:68)
        self.influent_mixed = influent_mod.sample_numpy(self._vec.scenario)
        obs = self._vec.reset(influent=torch.as_tensor(self.influent_mixed, dtype=torch.float64)[:, None])
        return obs[0].cpu().numpy()

    def step(self, action):
        if self.influent_mixed is None:
            raise RuntimeError("step() before reset()")
        action = np.clip(np.asarray(action, dtype=np.float64), self.action_space.low, self.action_space.high)
        self._last_action = action
        obs, reward, done, info = self._vec.step(torch.as_tensor(action, dtype=torch.float64)[None, :])
        self.reward = float(reward[0])
        self.info = {k: v[..., 0].cpu().numpy() for k, v in info.items()}
        return obs[0].cpu().numpy(), self.reward, True, {}

    def trajectory(self):
        """(t, x, kla) of the last step's cycle -- the first two return values of the reference's SBR_model_FB.run (`t`: list
        of times in days, `x`: [14, len(t)]; SBR_model_FB.py:71-86, 295) sampled at the END of every PID interval (the
        reference also lists the 8-9 interior output points of each), with the post-draw state at index 492, and the KLa
        of each interval (the reference's kla3 / kla5 / kla8 are kla[72:295], kla[481:492], kla[493:529])."""
        if getattr(self, "_last_action", None) is None:
            raise RuntimeError("trajectory() before step()")
        tr = self._vec.trajectory(torch.as_tensor(self._last_action, dtype=torch.float64)[None, :])
        return (tr["t"][:, 0].cpu().numpy().tolist(), tr["x"][:, :, 0].cpu().numpy().T.copy(),
                tr["kla"][:, 0].cpu().numpy())

    def render(self, mode="human", close=False):
        print("Reward for this episode: {}".format(self.reward))


class SbrEnv(_Base):
    """`SBR-v0` (gym_SBR_env0.py:139-262): one step = one whole 12-h cycle under the batch-to-batch (iterative-learning)
    feed-forward KLa plus the feedback DO-PID; action = DO set-points of phases 3, 5, 8 in [0, 5]; obs = the 14
    normalised sums x_last + influent (first entry 1); done = True after every step, the plant state carries over.

    The reference module runs cycle 0 at import and `reset()` only re-reads its globals (:150-176): here the first
    `reset()` runs cycle 0, later calls return the current observation without touching the plant.  Deviations, all
    forced by the reference (see gym_sbr2_b200/ilc.py): `step()` there raises before it returns (float linspace counts,
    seven-argument reward call), so the reward is module_reward.sbr_reward's formula on this cycle's applied KLa -- by
    construction, not pinned.  The influent is buffer_tank2.influent.buffer_tank(0, 12) on the global numpy RNG, bit-exact
    with the reference for the same np.random.seed (gym_SBR_env0.py:74,208).  learn="frozen" reproduces the module's behaviour (the controller keeps learning from cycle 0's
    memories, :200), learn="feedback" feeds the last cycle's memories back."""
    metadata = {"render.modes": ["human"]}

    def __init__(self, device=None, learn="frozen"):
        self.action_space = Box(np.array([0.0, 0.0, 0.0]), np.array([5.0, 5.0, 5.0]), dtype=np.float32)        # :145
        self.observation_space = Box(low=np.zeros(14), high=np.full(14, 2.0), dtype=np.float32)                # :147
        self.reward = 0
        self._vec = SbrIlcVecEnv(1, device=_device(device), learn=learn)
        self.influent_mixed = None
        self.info = {}

    def _draw(self):
        self.influent_mixed = influent_mod.sample_numpy_bt2()
        return torch.as_tensor(self.influent_mixed, dtype=torch.float64)[None, :]

    def reset(self):
        if self.influent_mixed is None:
            return self._vec.reset(influent=self._draw())[0].cpu().numpy()
        return self._vec._obs()[0].cpu().numpy()

    def step(self, action):
        if self.influent_mixed is None:
            raise RuntimeError("step() before reset()")
        action = np.clip(np.asarray(action, dtype=np.float64), ACTION_LOW, ACTION_HIGH)
        obs, reward, done, info = self._vec.step(torch.as_tensor(action, dtype=torch.float64)[None, :],
                                                 influent=self._draw())
        self.reward = float(reward[0])
        self.info = {k: (v[..., 0].cpu().numpy() if torch.is_tensor(v) else v) for k, v in info.items()}
        return obs[0].cpu().numpy(), self.reward, True, {}

    def render(self, mode="human", close=False):
        print("Reward for this episode: {}".format(self.reward))


class SbrEnv1(_Base):
    """`SBR-v1` (gym_SBR_env1.py:103-203): `SBR-v0`'s plant under its feedback DO-PID alone; one step = one whole cycle
    (SBR_model_FBc_implemented.run) from the state the previous step ended in; action = DO set-points of phases 3, 5, 8
    in [0, 5]; obs = the 14 normalised sums x + influent (first entry 1); done = True after every step.
    As in the reference, `reset()` never moves the plant: the first call returns the observation of the module's initial
    state, later calls that of the current one.  The reference's `step()` raises on a seven-argument call of the
    ten-parameter reward (:151); the reward here is that function's formula on the cycle's applied KLa (by construction)."""
    metadata = {"render.modes": ["human"]}

    def __init__(self, device=None):
        self.action_space = Box(np.array([0.0, 0.0, 0.0]), np.array([5.0, 5.0, 5.0]), dtype=np.float32)        # :109
        self.observation_space = Box(low=np.zeros(14), high=np.full(14, 2.0), dtype=np.float32)                # :111
        self.reward = 0
        self._vec = SbrV1VecEnv(1, device=_device(device))
        self.influent_mixed = None
        self.info = {}

    _draw = SbrEnv._draw

    def reset(self):
        if self.influent_mixed is None:
            return self._vec.reset(influent=self._draw())[0].cpu().numpy()
        return self._vec._obs()[0].cpu().numpy()

    step = SbrEnv.step
    render = SbrEnv.render


class SbrOS(_Base):
    """`SBROS-v1` (gym_SBR_oneshot.py:98-2644): one step = one 72-s PID interval; action = [DO set-point, NO3
    set-point]; returns the reference's 5-tuple (obs, state, reward, done, info) with obs = (obs_DO, obs_EC)."""
    metadata = {"render.modes": ["human"]}

    def __init__(self, device=None, mode="dp45", rtol=1e-8, atol=1e-10):
        self.action_space = Box(np.array([-1]), np.array([1]), dtype=np.float32)                           # :106
        self.observation_space = Box(low=np.array([0, 0, 0, -1, -1]), high=np.ones([5]) * 1.0, dtype=np.float32)
        self._vec = SbrOsVecEnv(1, device=_device(device), mode=mode, rtol=rtol, atol=atol, record_trajectory=True)
        self.influent_mixed = None
        self.reward = 0
        self._states = []

    def trajectory(self):
        """The reference's 18-tuple (gym_SBR_oneshot.py:1275-1288), same order:
        (t_t, x_t, u_DO_t, u_EC_t, state_t, So_t, Ss_t, EC, Sno_t, dcv_EC, ie_EC, e_EC, reward_t, reward_EQI_t,
        reward_OCI_t, reward_AE_t, reward_EC_t, Snh_t).  Time-resolved entries (t_t, x_t, So_t, Ss_t, EC, Sno_t,
        Snh_t) are sampled at the ENDS of the PID intervals (the reference also lists the 8-9 interior output points of
        every interval), starting with the state after the fill phase; u_DO_t / u_EC_t hold the clipped set-point in
        force in each interval; state_t and the reward lists have one entry per step().  The NO3-controller
        internals dcv_EC, ie_EC, e_EC are not recorded (None)."""
        tr = self._vec.trajectory(0)
        x = tr["x"]
        return (tr["t"].tolist(), x, tr["u_do"].tolist(), tr["u_ec"].tolist(), list(self._states), x[:, 8].tolist(),
                x[:, 2].tolist(), tr["ec"].tolist(), x[:, 9].tolist(), None, None, None, tr["reward"].tolist(),
                tr["reward_EQI"].tolist(), tr["reward_OCI"].tolist(), tr["reward_AE"].tolist(),
                tr["reward_EC"].tolist(), x[:, 10].tolist())

    def reset(self):
        self.influent_mixed = influent_mod.sample_numpy(self._vec.scenario)       # buffer_tank(6), :180
        obs_do, obs_ec = self._vec.reset(influent=torch.as_tensor(self.influent_mixed, dtype=torch.float64)[:, None])
        self._states = []
        return (obs_do[0].cpu().numpy().tolist(), obs_ec[0].cpu().numpy().tolist())

    def step(self, action):
        if self.influent_mixed is None:
            raise RuntimeError("step() before reset()")
        a = torch.as_tensor(np.asarray(action, dtype=np.float64).reshape(1, 2))
        (obs_do, obs_ec), state, reward, done, info = self._vec.step(a)
        self.reward = float(reward[0])
        obs = (obs_do[0].cpu().numpy().tolist(), obs_ec[0].cpu().numpy().tolist())
        state = state[0].cpu().numpy()
        self._states.append(state)
        return obs, state.tolist(), self.reward, bool(done[0]), {}

    def get_available_actions(self, pre_action, n_agents, n_action):
        """Action masks of the reference's discrete multi-agent wrapper (gym_SBR_oneshot.py:440-459)."""
        action_list = ([-0.1, 0, 0.1], [-5, 0, 5])
        action_boundary = ([0, 8], [0, 15])
        avail_us = []
        for agent_i in range(0, n_agents):
            avail_u = np.ones(n_action)
            for i in range(0, n_action):
                v = pre_action[agent_i] + action_list[agent_i][i]
                avail_u[i] = 1 if action_boundary[agent_i][0] <= v <= action_boundary[agent_i][1] else 0
            avail_us.append(avail_u)
        return avail_us

    def render(self, mode="human", close=False):
        print("Reward for this step: {}".format(self.reward))


class SbrEnv4(_Base):
    """`SBR-v4` (gym_SBR_env4.py:71-1294): one step = one 72-s PID interval incl. the fill phase; action = change of
    the DO set-point in [-1, 1]; obs = x / x_1 (14 values); standard 4-tuple.  Parity is against the reference run
    with numpy < 1.18 linspace semantics (its step() raises TypeError on current numpy) -- see SbrV4VecEnv."""
    metadata = {"render.modes": ["human"]}

    def __init__(self, device=None, mode="dp45", rtol=1e-8, atol=1e-10):
        self.action_space = Box(np.array([-1.0]), np.array([1.0]), dtype=np.float32)                       # :77
        self.observation_space = Box(low=0.9 * np.ones([14]), high=np.ones([14]), dtype=np.float32)        # :81
        self._vec = SbrV4VecEnv(1, device=_device(device), mode=mode, rtol=rtol, atol=atol)
        self.influent_mixed = None
        self.switch = None
        self.reward = 0

    def reset(self):
        self.switch, self.influent_mixed = influent_mod.sample_numpy_random_scenario()                      # :104
        obs = self._vec.reset(influent=torch.as_tensor(self.influent_mixed, dtype=torch.float64)[:, None])
        return obs.cpu().numpy()                                     # shape (1, 14) like the reference's x_2 / x_1

    def step(self, action):
        if self.influent_mixed is None:
            raise RuntimeError("step() before reset()")
        a = torch.as_tensor(np.asarray(action, dtype=np.float64).reshape(1))
        obs, reward, done, info = self._vec.step(a)
        self.reward = float(reward[0])
        return obs[0].cpu().numpy(), self.reward, bool(done[0]), {}

    def render(self, mode="human", close=False):
        print("Reward for this step: {}".format(self.reward))


class _CntEnv(_Base):
    """Batch-of-one view of SbrCntVecEnv under the reference's class name.  The reference's step() of these ids
    raises NameError inside module_reward_continuous1.sbr_reward; states / observations / done follow the unmodified
    env modules, the reward is the repaired form disclosed in oracle/make_golden_cnt.py."""
    metadata = {"render.modes": ["human"]}
    kind = None

    def __init__(self, device=None, mode="dp45", rtol=1e-8, atol=1e-10):
        self._vec = SbrCntVecEnv(self.kind, 1, device=_device(device), mode=mode, rtol=rtol, atol=atol)
        self.influent_mixed = None
        self.reward = 0
        self._log = dict(t=[], x=[], u_do=[], u_ec=[], state=[], ec=[])

    def _record(self, state=None):
        st = self._vec.buf.st[:, 0].cpu().numpy()
        L = self._log
        L["t"].append(float(st[14])); L["x"].append(st[:14].copy()); L["u_do"].append(float(st[15]))
        L["u_ec"].append(float(st[16])); L["ec"].append(float(st[23]))
        if state is not None:
            L["state"].append(np.asarray(state, dtype=np.float64).reshape(-1))

    def trajectory(self):
        """What the reference's trajectory() hands a plotting script (gym_SBR_continuous1.py:423-436: t_t, x_t, u_t,
        state_t, So_t; the carbon-controller envs add u_EC_t, Ss_t, (Sno_t,) EC: gym_SBR_continuous2.py:483-496,
        gym_SBR_continuous_MA1.py:506-519), sampled at the END of every step() (the reference also lists the 8-9
        interior output points of every control interval), starting with the state after the fill phase.  Returns a dict
        with the reference's names: t_t, x_t [T,14], u_DO_t (`u_t` of SbrCnt1), u_EC_t, state_t, So_t, Ss_t, Sno_t, EC."""
        L = self._log
        x = np.array(L["x"]) if L["x"] else np.zeros((0, 14))
        return dict(t_t=list(L["t"]), x_t=x, u_DO_t=list(L["u_do"]), u_EC_t=list(L["u_ec"]), state_t=list(L["state"]),
                    So_t=x[:, 8].tolist(), Ss_t=x[:, 2].tolist(), Sno_t=x[:, 9].tolist(), EC=list(L["ec"]))

    def _shape_obs(self, obs):
        return obs[0].cpu().numpy()

    def reset(self):
        self.influent_mixed = influent_mod.sample_numpy(self._vec.scenario)         # buffer_tank(0) in all five files
        obs = self._vec.reset(influent=torch.as_tensor(self.influent_mixed, dtype=torch.float64)[:, None])
        self._log = dict(t=[], x=[], u_do=[], u_ec=[], state=[], ec=[])
        out = self._shape_reset(obs)
        self._record(None if self.kind == "os2" else out)
        return out

    def _shape_reset(self, obs):
        return self._shape_obs(obs)

    def step(self, action):
        if self.influent_mixed is None:
            raise RuntimeError("step() before reset()")
        a = torch.as_tensor(np.asarray(action, dtype=np.float64).reshape(1))
        obs, reward, done, info = self._vec.step(a)
        self.reward = float(reward[0])
        out = self._shape_obs(obs)
        self._record(out)
        return out, self.reward, bool(done[0]), {}

    def render(self, mode="human", close=False):
        print("Reward for this step: {}".format(self.reward))


class SbrCnt0(_CntEnv):
    """`SBRCnt-v0` (gym_SBR_continuous0.py:85-1284): action = change of the DO set-point (declared [-0.05, 0.05]);
    state = [t, Si, Xbh, Xba, So, Sno, Snh] / x_1, shape (1, 7) like the reference's x_2 / x_1; 466 steps per episode."""
    kind = "cnt0"

    def __init__(self, **kw):
        self.action_space = Box(np.array([-0.05]), np.array([0.05]), dtype=np.float32)                     # :92
        self.observation_space = Box(low=np.array([0, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9]), high=np.ones([7]) * 1.3,
                                     dtype=np.float32)                                                     # :96-98
        super().__init__(**kw)

    def _shape_obs(self, obs):
        return obs.cpu().numpy()                                     # (1, 7)


class SbrCnt1(_CntEnv):
    """`SBRCnt-v1` (gym_SBR_continuous1.py:82-1399): the agent moves the DO set-point in the aerobic phases only (228
    steps per episode; the anoxic phases are simulated whole inside a step); state = [t/0.5, So/8, Snh/30, dSo, dSnh]."""
    kind = "cnt1"

    def __init__(self, **kw):
        self.action_space = Box(np.array([-1]), np.array([1]), dtype=np.float32)                           # :89
        self.observation_space = Box(low=np.array([0, 0, 0, -1, -1]), high=np.ones([5]) * 1.0, dtype=np.float32)
        super().__init__(**kw)


class SbrCnt2(SbrCnt1):
    """`SBRCnt-v2` (gym_SBR_continuous2.py:96-1560): SbrCnt1 plus a carbon controller on Ss whose set-point the first
    action also moves."""
    kind = "cnt2"


class SbrCntMA1(SbrCnt1):
    """`SBRCntMA-v1` (gym_SBR_continuous_MA1.py:96-1578): one 1-D action moves the carbon set-point (on Sno) in the
    anoxic phases and the DO set-point in the aerobic ones; 463 steps per episode."""
    kind = "ma1"


class SbrOS1(_CntEnv):
    """`SBROS-v2` (gym_SBR_oneshot1.py:98-2089): absolute [DO set-point, NO3 set-point] actions; returns the SBROS-v1
    style 5-tuple (obs, state, reward, done, info) with obs = (obs_DO, obs_EC)."""
    kind = "os2"

    def __init__(self, **kw):
        self.action_space = Box(np.array([-1]), np.array([1]), dtype=np.float32)                           # :105
        self.observation_space = Box(low=np.array([0, 0, 0, -1, -1]), high=np.ones([5]) * 1.0, dtype=np.float32)
        super().__init__(**kw)

    def _shape_reset(self, obs):
        return (obs[0][0].cpu().numpy().tolist(), obs[1][0].cpu().numpy().tolist())

    def step(self, action):
        if self.influent_mixed is None:
            raise RuntimeError("step() before reset()")
        a = torch.as_tensor(np.asarray(action, dtype=np.float64).reshape(1, 2))
        (obs_do, obs_ec), state, reward, done, info = self._vec.step(a)
        self.reward = float(reward[0])
        obs = (obs_do[0].cpu().numpy().tolist(), obs_ec[0].cpu().numpy().tolist())
        state = state[0].cpu().numpy()
        self._record(state)
        return obs, state, self.reward, bool(done[0]), {}

    def get_available_actions(self, pre_action, n_agents, n_action):
        """Action masks of the discrete multi-agent wrapper (gym_SBR_oneshot1.py:434-454)."""
        action_list = ([-1, -0.5, 0, 0.5, 1], [-1, -0.5, 0, 0.5, 1])
        action_boundary = ([0, 8], [0, 15])
        avail_us = []
        for agent_i in range(0, n_agents):
            avail_u = np.ones(n_action)
            for i in range(0, n_action):
                v = pre_action[agent_i] + action_list[agent_i][i]
                avail_u[i] = 1 if action_boundary[agent_i][0] <= v <= action_boundary[agent_i][1] else 0
            avail_us.append(avail_u)
        return avail_us
