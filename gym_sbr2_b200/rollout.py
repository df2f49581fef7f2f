"""PPO-style rollouts over the vector envs (BASELINE config 5): whole episodes stepped on the device, observations
and rewards staying in torch tensors, per-env episode returns gathered over NCCL at the end -- the only collective.

The policy here is a stand-in (a tiny fixed MLP from the 18 observation values to the two set-points); what is
exercised is the data path an RL trainer uses: obs tensors -> policy -> action tensor -> one kernel launch per step.
"""
import torch

from . import dist


class TinyPolicy(torch.nn.Module):
    """obs_DO (9) ++ obs_EC (9) -> [DO set-point in (0.5, 7), NO3 set-point in (1, 14)] (float32 MLP, fixed seed)."""

    def __init__(self, device, seed=0, hidden=32):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w1 = (torch.randn(18, hidden, generator=g) * 0.3).to(device)
        self.w2 = (torch.randn(hidden, 2, generator=g) * 0.3).to(device)
        self.lo = torch.tensor([0.5, 1.0], device=device)
        self.span = torch.tensor([6.5, 13.0], device=device)

    @torch.no_grad()
    def forward(self, obs_do, obs_ec):
        x = torch.cat([obs_do, obs_ec], dim=1).to(torch.float32)
        y = torch.sigmoid(torch.tanh(x @ self.w1) @ self.w2)
        return (self.lo + self.span * y).to(torch.float64)


@torch.no_grad()
def collect_episode(env, policy, max_steps=None, store=False):
    """Run one SBROS-v1 episode for every env of `env` (a SbrOsVecEnv).  Returns dict(returns [N], steps, and, with
    store=True, the [T,N] reward / done buffers a PPO update would consume)."""
    obs_do, obs_ec = env.reset()
    n = env.num_envs
    steps = max_steps or env.max_episode_steps
    rewards = torch.empty((steps, n), dtype=torch.float64, device=env.device) if store else None
    dones = torch.empty((steps, n), dtype=torch.bool, device=env.device) if store else None
    k = 0
    for k in range(steps):
        action = policy(obs_do, obs_ec)
        (obs_do, obs_ec), state, reward, done, info = env.step(action)
        if store:
            rewards[k].copy_(reward)
            dones[k].copy_(done)
    return dict(returns=info["episode_return"].clone(), steps=k + 1, rewards=rewards, dones=dones,
                all_done=done.all(), status=info["status"])


def return_stats(allr):
    ok = torch.isfinite(allr)
    r = allr[ok]
    return dict(mean=float(r.mean()), std=float(r.std(unbiased=False)), min=float(r.min()), max=float(r.max()),
                count=int(ok.sum()))


def gather_episode_returns(local_returns, n_total):
    """NCCL all_gather of per-env episode returns across ranks (global env order), plus the 5 summary statistics."""
    allr = dist.gather_rewards(local_returns, n_total)
    return allr, return_stats(allr)
