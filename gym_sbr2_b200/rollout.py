"""PPO-style rollouts over the vector envs (BASELINE config 5): whole episodes stepped on the device, observations
and rewards staying in torch tensors, per-env episode returns gathered over NCCL at the end -- the only collective.

The policy here is a stand-in (a tiny fixed MLP from the 18 observation values to the two set-points); what is
exercised is the data path an RL trainer uses: obs tensors -> policy -> action tensor -> one kernel launch per step.
"""
import torch

from . import core, dist


class TinyPolicy(torch.nn.Module):
    """obs_DO (9) ++ obs_EC (9) -> [DO set-point in (0.5, 7), NO3 set-point in (1, 14)] (float32 MLP, fixed seed).
    Works directly on the kernels' struct-of-arrays layout: input [18, N], output [2, N] -- no transposes.
    n_in / lo / span change the shape: TinyPolicy(device, n_in=14, lo=[-0.2], span=[0.4]) is the SBR-v4 stand-in
    (14 observations -> one change of the DO set-point)."""

    def __init__(self, device, seed=0, hidden=32, n_in=18, lo=(0.5, 1.0), span=(6.5, 13.0)):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        n_out = len(lo)
        self.n_in, self.n_out = int(n_in), n_out
        self.w1t = (torch.randn(hidden, n_in, generator=g) * 0.3).to(device)
        self.w2t = (torch.randn(n_out, hidden, generator=g) * 0.3).to(device)
        self.lo = torch.tensor([[v] for v in lo], dtype=torch.float32, device=device)
        self.span = torch.tensor([[v] for v in span], dtype=torch.float32, device=device)
        self.w1t, self.w2t = self.w1t.contiguous(), self.w2t.contiguous()
        self.lo_flat, self.span_flat = self.lo.reshape(-1).contiguous(), self.span.reshape(-1).contiguous()

    @torch.no_grad()
    def forward_torch(self, obs_do, obs_ec):
        """The policy as a plain torch expression (nine launches, ~1.3 GB of traffic per step at 2^20 envs): the
        fp32 reference the fused kernel is tested against."""
        x = (obs_do if obs_ec is None else torch.cat([obs_do, obs_ec], dim=0)).to(torch.float32)
        y = torch.sigmoid(self.w2t @ torch.tanh(self.w1t @ x))
        return (self.lo + self.span * y).to(torch.float64)

    @torch.no_grad()
    def act_into(self, obs_do, obs_ec, action):
        """obs_do, obs_ec: [9, N] float64 -> action [2, N] float64, ONE launch (sbr_policy_mlp)."""
        return core.policy_mlp(obs_do, obs_ec, self.w1t, self.w2t, self.lo_flat, self.span_flat, action)

    def as_struct(self):
        """The weights as the C ABI's SbrPolicyMlp (for the fused rollout, sbr_os_rollout_k)."""
        from . import _abi
        return _abi.SbrPolicyMlp(self.w1t.data_ptr(), self.w2t.data_ptr(), self.lo_flat.data_ptr(),
                                 self.span_flat.data_ptr(), self.n_in, self.w1t.shape[0], self.n_out, 0)

    def forward_soa(self, obs_do, obs_ec):
        """obs_do, obs_ec: [9, N] float64 -> action [2, N] float64 (a fresh tensor)."""
        out = torch.empty((2, obs_do.shape[1]), dtype=torch.float64, device=obs_do.device)
        return self.act_into(obs_do, obs_ec, out)

    def forward(self, obs_do, obs_ec):
        """Gym layout: obs [N, 9] each -> action [N, 2]."""
        return self.forward_soa(obs_do.t(), obs_ec.t()).t()


@torch.no_grad()
def collect_episode(env, policy, max_steps=None, store=False):
    """Run one SBROS-v1 episode for every env of `env` (a SbrOsVecEnv), observations and actions staying in the
    kernels' SoA layout.  Returns dict(returns [N], steps, and, with store=True, the [T,N] reward / done buffers a
    PPO update would consume)."""
    env.reset()
    b = env.buf
    n = env.num_envs
    steps = max_steps or env.max_episode_steps
    rewards = torch.empty((steps, n), dtype=torch.float64, device=env.device) if store else None
    dones = torch.empty((steps, n), dtype=torch.bool, device=env.device) if store else None
    k = 0
    for k in range(steps):
        action = policy.act_into(b.obs_do, b.obs_ec, env._action)
        env.step_soa(action)
        if store:
            rewards[k].copy_(b.reward)
            dones[k].copy_(b.done)
    from . import _abi
    return dict(returns=b.st[_abi.OS_RETURN].clone(), steps=k + 1, rewards=rewards, dones=dones,
                all_done=b.done.bool().all(), status=b.status)


class GraphedStepper(object):
    """CUDA-graph version of the rollout inner loop for the launch-bound regime (BASELINE config[1]: 4096 envs, where
    one env.step is ~10 us of GPU work behind ~160 us of Python + launches): `steps_per_replay` iterations of
    [policy on the SoA observation buffers -> action, sbr_os_step] are captured once on the env's static buffers
    and replayed with a single launch.  The C ABI launches on the stream it is given (torch's current stream), so
    capture needs nothing special.  Episode ends are not handled inside the graph: replay whole episodes
    (463 = 57 x 8 + 7 steps) and reset between them, or use steps_per_replay=1."""

    def __init__(self, env, policy, steps_per_replay=8, warmup=3):
        self.env, self.policy, self.k = env, policy, int(steps_per_replay)
        self.action = torch.zeros((2, env.num_envs), dtype=torch.float64, device=env.device)
        b = env.buf
        side = torch.cuda.Stream(device=env.device)
        side.wait_stream(torch.cuda.current_stream(env.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self._one(b)
        torch.cuda.current_stream(env.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            for _ in range(self.k):
                self._one(b)
        # replays advance the envs without passing through the env's host-side episode clock: autoreset must not
        # rely on it any more (it falls back to its masked, every-step path; a full reset() re-arms the clock, and
        # replay() below takes it out again)
        env._lockstep = False

    def _one(self, b):
        self.policy.act_into(b.obs_do, b.obs_ec, self.action)
        self.env.step_soa(self.action)

    def replay(self):
        """Advance every env by steps_per_replay env.steps."""
        self.env._lockstep = False
        self.graph.replay()


@torch.no_grad()
def collect_episode_graphed(env, big, small):
    """One whole episode through two captured graphs (`big`: k steps per replay, `small`: 1 step per replay)."""
    env.reset()
    steps = env.max_episode_steps
    for _ in range(steps // big.k):
        big.replay()
    for _ in range(steps % big.k):
        small.replay()
    from . import _abi
    b = env.buf
    return dict(returns=b.st[_abi.OS_RETURN].clone(), steps=steps, all_done=b.done.bool().all(), status=b.status)


@torch.no_grad()
def collect_episode_fused(env, policy, K=8, store=False):
    """One whole SBROS-v1 episode through the FUSED rollout kernel (sbr_os_rollout_k): K env.steps per launch with the
    policy evaluated in-kernel between them -- observations stay on the SM, one launch moves the state once per K steps.
    Bit-identical to collect_episode (same step and policy arithmetic).  store=True also returns the [T,N] rewards and
    the per-step set-points / observations a PPO update consumes (written as streams by the kernel)."""
    from . import _abi
    env.reset()
    b = env.buf
    n = env.num_envs
    steps = env.max_episode_steps
    pol = policy.as_struct()
    action = env._action
    policy.act_into(b.obs_do, b.obs_ec, action)                       # the first step's set-points, from the reset observation
    nl = (steps + K - 1) // K
    rewards = torch.zeros((nl * K, n), dtype=torch.float64, device=env.device)
    acts = torch.empty((nl * K, 2, n), dtype=torch.float64, device=env.device) if store else None
    obs = torch.empty((nl * K, 18, n), dtype=torch.float64, device=env.device) if store else None
    env._lockstep = False
    done_steps = 0
    for j in range(nl):
        k = min(K, steps - done_steps)
        sl = slice(j * K, j * K + k)
        core.os_rollout_k(b, action, pol, rewards[sl], env.params, env.sched, mode=env.mode, tol=env.tol,
                          emit=env.emit, act_log=None if acts is None else acts[sl],
                          obs_log=None if obs is None else obs[sl])
        done_steps += k
    return dict(returns=b.st[_abi.OS_RETURN].clone(), steps=steps, all_done=b.done.bool().all(), status=b.status,
                rewards=rewards[:steps] if store else None, actions=None if acts is None else acts[:steps],
                observations=None if obs is None else obs[:steps])


@torch.no_grad()
def collect_episode_v4(env, policy, max_steps=None):
    """One SBR-v4 episode step by step: [sbr_policy_mlp on obs = x / x_1 -> change of the DO set-point, sbr_v4_step]."""
    from . import _abi
    env.reset()
    b = env.buf
    steps = max_steps or env.max_episode_steps
    act = torch.empty((1, env.num_envs), dtype=torch.float64, device=env.device)
    for _ in range(steps):
        policy.act_into(b.obs, None, act)
        env.step_async(act[0])
    return dict(returns=env._st_row(_abi.V4_RETURN).clone(), steps=steps, all_done=b.done.bool().all(), status=b.status)


@torch.no_grad()
def collect_episode_v4_fused(env, policy, K=8, resort_every=2):
    """One whole SBR-v4 episode through the fused rollout kernel (sbr_v4_rollout_k): K steps per launch with the policy head
    evaluated in-kernel.  Every device buffer of the rollout lives in SLOT order -- slots fully re-sorted (single envs) by
    the last launch's RHS count every `resort_every` launches with one sbr_permute_rows -- so the kernel has no
    env-indexed access at all; the slot -> env map is only applied to what leaves the rollout (the episode returns).
    Bit-identical to collect_episode_v4 (results do not depend on the placement)."""
    from . import _abi
    env.unsort()                                      # the collector manages the placement itself
    env.reset()
    b = env.buf
    n, dev = env.num_envs, env.device
    steps = env.max_episode_steps
    pol = policy.as_struct()
    action = torch.empty((1, n), dtype=torch.float64, device=dev)
    policy.act_into(b.obs, None, action)              # the first step's action, from the reset observation
    action = action[0].contiguous()
    load = env._loading
    nl = (steps + K - 1) // K
    rewards = torch.zeros((K, n), dtype=torch.float64, device=dev)
    slot_env = None                                   # int64 [n]: slot j holds env slot_env[j]
    alt = dict(st=torch.empty_like(b.st), load=torch.empty_like(load), action=torch.empty_like(action),
               done=torch.empty((n,), dtype=torch.int32, device=dev))
    done_steps = 0
    for j in range(nl):
        k = min(K, steps - done_steps)
        core.v4_rollout_k(b, load, action, pol, rewards[:k], env.params, env.sched, mode=env.mode, tol=env.tol,
                          emit_obs=False)
        done_steps += k
        if resort_every and (j + 1) % resort_every == 0 and j + 1 < nl:
            perm = torch.argsort(b.counters[0])
            done32 = b.done.to(torch.int32)
            core.permute_rows(perm, [(b.st, alt["st"]), (load, alt["load"]), (action, alt["action"]), (done32, alt["done"])])
            b.st, alt["st"] = alt["st"], b.st
            load, alt["load"] = alt["load"], load
            action, alt["action"] = alt["action"], action
            b.done.copy_(alt["done"].to(torch.uint8))
            slot_env = perm if slot_env is None else slot_env[perm]
    ret, all_done, status = b.st[_abi.V4_RETURN].clone(), b.done.bool().all(), b.status
    if slot_env is not None:                           # back to env order: state, returns
        out = torch.empty_like(ret)
        out[slot_env] = ret
        ret = out
        core.permute_rows(slot_env, [(b.st, alt["st"])], scatter=True)
        b.st, alt["st"] = alt["st"], b.st
    env._lockstep = False
    return dict(returns=ret, steps=steps, all_done=all_done, status=status)


@torch.no_grad()
def collect_episode_v4_sorted(env, policy, resort_every=8):
    """One SBR-v4 episode STEP BY STEP (one policy evaluation and one sbr_v4_step launch per env.step, for policies that
    cannot move into the fused kernel) with every device buffer of the rollout in SLOT order, as collect_episode_v4_fused
    keeps them: slots fully re-sorted (single envs) by the last step's RHS count every `resort_every` steps with one
    sbr_permute_rows of state, loading, observation and done flags.  A policy that maps each env's observation to that
    env's action does not care in which order the envs come; the slot -> env map is applied to the returns at the end.
    `policy` needs act_into(obs, None, action) (TinyPolicy) -- any per-env torch expression on obs [14, N] works the same
    way.  Bit-identical to collect_episode_v4."""
    from . import _abi
    env.unsort()
    env.reset()
    b = env.buf
    n, dev = env.num_envs, env.device
    steps = env.max_episode_steps
    act = torch.empty((1, n), dtype=torch.float64, device=dev)
    load = env._loading
    slot_env = None
    alt = dict(st=torch.empty_like(b.st), load=torch.empty_like(load), obs=torch.empty_like(b.obs),
               done=torch.empty((n,), dtype=torch.int32, device=dev))
    for k in range(steps):
        policy.act_into(b.obs, None, act)
        core.v4_step(b, load, act[0], env.params, env.sched, mode=env.mode, tol=env.tol)
        if resort_every and (k + 1) % resort_every == 0 and k + 1 < steps:
            perm = torch.argsort(b.counters[0])
            done32 = b.done.to(torch.int32)
            core.permute_rows(perm, [(b.st, alt["st"]), (load, alt["load"]), (b.obs, alt["obs"]), (done32, alt["done"])])
            b.st, alt["st"] = alt["st"], b.st
            load, alt["load"] = alt["load"], load
            b.obs, alt["obs"] = alt["obs"], b.obs
            b.done.copy_(alt["done"].to(torch.uint8))
            slot_env = perm if slot_env is None else slot_env[perm]
    ret, all_done, status = b.st[_abi.V4_RETURN].clone(), b.done.bool().all(), b.status
    if slot_env is not None:
        out = torch.empty_like(ret)
        out[slot_env] = ret
        ret = out
        core.permute_rows(slot_env, [(b.st, alt["st"]), (b.obs, alt["obs"])], scatter=True)
        b.st, alt["st"] = alt["st"], b.st
        b.obs, alt["obs"] = alt["obs"], b.obs
    env._lockstep = False
    return dict(returns=ret, steps=steps, all_done=all_done, status=status)


def _cnt_policy_obs(env):
    """The observation rows the policy head of a SBRCnt / SBROS-v2 env reads (all of them; os2: obs_DO ++ obs_EC)."""
    o = env.buf.obs
    return o[:18] if env.kind == "os2" else o


@torch.no_grad()
def collect_episode_cnt(env, policy, max_steps=None):
    """One episode of a SbrCntVecEnv step by step: [sbr_policy_mlp on the env's observation rows, sbr_cnt_step]."""
    from . import _abi, cnt
    env.reset()
    b = env.buf
    steps = max_steps or env.max_episode_steps
    for _ in range(steps):
        a = env._action if env.kind == "os2" else env._action[:1]
        policy.act_into(_cnt_policy_obs(env), None, a)
        cnt.cnt_step(env.cfg, b, env._action, env.params, env.sched, mode=env.mode, tol=env.tol)
    return dict(returns=b.st[_abi.CNT_RETURN].clone(), steps=steps, all_done=b.done.bool().all(), status=b.status)


@torch.no_grad()
def collect_episode_cnt_fused(env, policy, K=8):
    """The same episode through the fused rollout kernel (sbr_cnt_rollout_k): K steps per launch, policy head in-kernel."""
    from . import _abi, cnt
    env.reset()
    b = env.buf
    n = env.num_envs
    steps = env.max_episode_steps
    pol = policy.as_struct()
    a = env._action if env.kind == "os2" else env._action[:1]
    policy.act_into(_cnt_policy_obs(env), None, a)
    rewards = torch.zeros((K, n), dtype=torch.float64, device=env.device)
    done_steps = 0
    while done_steps < steps:
        k = min(K, steps - done_steps)
        cnt.cnt_rollout_k(env.cfg, b, env._action, pol, rewards[:k], env.params, env.sched, mode=env.mode, tol=env.tol)
        done_steps += k
    return dict(returns=b.st[_abi.CNT_RETURN].clone(), steps=steps, all_done=b.done.bool().all(), status=b.status)


def return_stats(allr):
    ok = torch.isfinite(allr)
    r = allr[ok]
    return dict(mean=float(r.mean()), std=float(r.std(unbiased=False)), min=float(r.min()), max=float(r.max()),
                count=int(ok.sum()))


def gather_episode_returns(local_returns, n_total):
    """NCCL all_gather of per-env episode returns across ranks (global env order), plus the 5 summary statistics."""
    allr = dist.gather_rewards(local_returns, n_total)
    return allr, return_stats(allr)
