"""PPO-style rollouts over the vector envs (BASELINE config 5): whole episodes stepped on the device, observations
and rewards staying in torch tensors, per-env episode returns gathered over NCCL at the end -- the only collective.

The policy here is a stand-in (a tiny fixed MLP from the 18 observation values to the two set-points); what is
exercised is the data path an RL trainer uses: obs tensors -> policy -> action tensor -> one kernel launch per step.
"""
import torch

from . import dist


class TinyPolicy(torch.nn.Module):
    """obs_DO (9) ++ obs_EC (9) -> [DO set-point in (0.5, 7), NO3 set-point in (1, 14)] (float32 MLP, fixed seed).
    Works directly on the kernels' struct-of-arrays layout: input [18, N], output [2, N] -- no transposes."""

    def __init__(self, device, seed=0, hidden=32):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w1t = (torch.randn(hidden, 18, generator=g) * 0.3).to(device)
        self.w2t = (torch.randn(2, hidden, generator=g) * 0.3).to(device)
        self.lo = torch.tensor([[0.5], [1.0]], device=device)
        self.span = torch.tensor([[6.5], [13.0]], device=device)

    @torch.no_grad()
    def forward_soa(self, obs_do, obs_ec):
        """obs_do, obs_ec: [9, N] float64 -> action [2, N] float64."""
        x = torch.cat([obs_do, obs_ec], dim=0).to(torch.float32)
        y = torch.sigmoid(self.w2t @ torch.tanh(self.w1t @ x))
        return (self.lo + self.span * y).to(torch.float64)

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
        action = policy.forward_soa(b.obs_do, b.obs_ec)
        env.step_soa(action.contiguous())
        if store:
            rewards[k].copy_(b.reward)
            dones[k].copy_(b.done)
    from . import _abi
    return dict(returns=b.st[_abi.OS_RETURN].clone(), steps=k + 1, rewards=rewards, dones=dones,
                all_done=b.done.bool().all(), status=b.status)


def return_stats(allr):
    ok = torch.isfinite(allr)
    r = allr[ok]
    return dict(mean=float(r.mean()), std=float(r.std(unbiased=False)), min=float(r.min()), max=float(r.max()),
                count=int(ok.sum()))


def gather_episode_returns(local_returns, n_total):
    """NCCL all_gather of per-env episode returns across ranks (global env order), plus the 5 summary statistics."""
    allr = dist.gather_rewards(local_returns, n_total)
    return allr, return_stats(allr)
