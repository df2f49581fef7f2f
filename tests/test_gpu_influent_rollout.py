"""The step before the hot path (influent generator, buffer_tank3.py) on the device, and the PPO-style rollout
helper (BASELINE config 5) on one GPU."""
import numpy as np
import pytest
import torch

from gym_sbr2_b200 import _abi, core, influent, rollout
from gym_sbr2_b200.vec_env import SbrOsVecEnv, SbrV2VecEnv, SbrV4VecEnv

pytestmark = pytest.mark.gpu


def test_influent_kernel_bit_exact_with_numpy_for_all_scenarios(built, cuda_device):
    rng = np.random.RandomState(3)
    n = 777                                                       # ragged: not a multiple of the block
    rnd = rng.randn(48, n)
    for sw in range(8):
        got = core.influent_mix(sw, torch.as_tensor(rnd).to(cuda_device)).cpu().numpy()
        ref = np.stack([influent.mix_numpy(sw, rnd[:, i]) for i in range(n)], axis=1)
        assert np.array_equal(got, ref), sw


def test_influent_kernel_reproduces_reference_draws(built, cuda_device, golden_v2):
    """Same numpy stream as the reference (np.random.seed(s); buffer_tank(0)) -> the reference's influent_mixed,
    bit for bit (golden fixtures were recorded from the unmodified reference)."""
    g = golden_v2
    seeds = sorted(set(int(s) for s in g["seed"]))
    rnd = np.stack([np.random.RandomState(s).randn(48) for s in seeds], axis=1)
    got = core.influent_mix(0, torch.as_tensor(rnd).to(cuda_device)).cpu().numpy()
    for j, s in enumerate(seeds):
        i = int(np.nonzero(g["seed"] == s)[0][0])
        assert np.array_equal(got[:, j], g["influent"][i]), s


def test_vec_env_numpy_rng_matches_sequential_reference_resets(built, cuda_device):
    """rng='numpy': N envs consume the numpy stream like N sequential reference resets would (scenario 6 draws two
    randn(48) per reset and uses the second, buffer_tank3.py:206,224)."""
    n = 5
    env = SbrOsVecEnv(n, device=cuda_device, seed=11, rng="numpy")
    env.reset()
    r = np.random.RandomState(11)
    ref = np.stack([influent.sample_numpy(6, r) for _ in range(n)], axis=1)
    assert np.array_equal(env.influent.cpu().numpy(), ref)
    env2 = SbrV2VecEnv(n, device=cuda_device, seed=11, rng="numpy")
    env2.reset()
    r = np.random.RandomState(11)
    ref2 = np.stack([influent.sample_numpy(0, r) for _ in range(n)], axis=1)
    assert np.array_equal(env2.influent.cpu().numpy(), ref2)


def test_rollout_collects_full_episodes(built, cuda_device):
    n = 512
    env = SbrOsVecEnv(n, device=cuda_device, seed=5, mode="dp45")
    policy = rollout.TinyPolicy(cuda_device)
    ep = rollout.collect_episode(env, policy, store=True)
    assert ep["steps"] == 463 and bool(ep["all_done"])
    assert ep["rewards"].shape == (463, n) and bool(torch.isfinite(ep["rewards"]).all())
    assert bool(ep["dones"][-1].all()) and not bool(ep["dones"][:-1].any())
    assert torch.allclose(ep["returns"], ep["rewards"].sum(dim=0), rtol=1e-12, atol=1e-13)
    allr, stats = rollout.gather_episode_returns(ep["returns"], n)           # world size 1: identity
    assert torch.equal(allr, ep["returns"]) and stats["count"] == n
    assert int(ep["status"].max()) == 0


def test_graphed_rollout_matches_eager(built, cuda_device):
    """The CUDA-graph inner loop (policy + step captured, 8 steps per replay) reproduces the eager rollout bit for bit."""
    n = 256
    policy = rollout.TinyPolicy(cuda_device)
    env_e = SbrOsVecEnv(n, device=cuda_device, seed=5, mode="dp45")
    eager = rollout.collect_episode(env_e, policy)
    env_g = SbrOsVecEnv(n, device=cuda_device, seed=5, mode="dp45")
    env_g.reset()
    big, small = rollout.GraphedStepper(env_g, policy, 8), rollout.GraphedStepper(env_g, policy, 1)
    env_g.epoch.zero_()                                          # same influent draw as the eager env's first reset
    ep = rollout.collect_episode_graphed(env_g, big, small)
    assert bool(ep["all_done"]) and ep["steps"] == 463
    assert torch.equal(ep["returns"], eager["returns"])
    assert torch.equal(env_g.buf.st[:14], env_e.buf.st[:14])


@pytest.mark.parametrize("cls,nact", [(SbrOsVecEnv, 2), (SbrV4VecEnv, 1)])
def test_checkpoint_resume_mid_episode(built, cuda_device, tmp_path, cls, nact):
    """torch.save(env.state_dict()) mid-episode, restore into a fresh env, continue: bit-identical to the
    uninterrupted run (the reference has no resume path; its state lives in module globals)."""
    n = 128
    g = torch.Generator(device=cuda_device).manual_seed(3)
    acts = [torch.rand((n, nact), dtype=torch.float64, device=cuda_device, generator=g) * (4 if nact == 2 else 0.1)
            for _ in range(80)]
    a = cls(n, device=cuda_device, seed=9)
    a.reset()
    for k in range(40):
        a.step(acts[k])
    path = str(tmp_path / "ckpt.pt")
    torch.save(a.state_dict(), path)
    for k in range(40, 80):
        out_a = a.step(acts[k])
    b = cls(n, device=cuda_device, seed=1234)                      # different seed: everything comes from the file
    b.load_state_dict(torch.load(path, weights_only=False))
    for k in range(40, 80):
        out_b = b.step(acts[k])
    assert torch.equal(a.buf.st[:20], b.buf.st[:20])
    ra, rb = (out_a[2], out_b[2]) if nact == 2 else (out_a[1], out_b[1])
    assert torch.equal(ra, rb)
    assert torch.equal(a.reset()[0] if nact == 2 else a.reset(), b.reset()[0] if nact == 2 else b.reset())   # RNG too


@pytest.mark.parametrize("rng", ["philox", "numpy"])
def test_checkpoint_resume_policy_driven(built, cuda_device, tmp_path, rng):
    """Resume with the policy computing its first action from the RESTORED observation buffers (they are part of the
    checkpoint, not torch.empty of a fresh env), across an autoreset so that the restored RNG state is used too."""
    n = 96
    policy = rollout.TinyPolicy(cuda_device)

    def advance(env, k):
        for _ in range(k):
            b = env.buf
            out = env.step(policy.forward(b.obs_do.t(), b.obs_ec.t()))
        return out

    a = SbrOsVecEnv(n, device=cuda_device, seed=21, rng=rng, autoreset=True)
    a.reset()
    advance(a, 430)
    path = str(tmp_path / "ckpt.pt")
    torch.save(a.state_dict(), path)
    out_a = advance(a, 60)                                                   # crosses the episode end at step 463
    b = SbrOsVecEnv(n, device=cuda_device, seed=5, rng="philox", autoreset=True)
    b.load_state_dict(torch.load(path, weights_only=False))
    out_b = advance(b, 60)
    assert torch.equal(a.buf.st.view(torch.int64), b.buf.st.view(torch.int64))      # bit patterns: Qw is NaN mid-episode
    assert torch.equal(a.influent, b.influent)
    assert torch.equal(out_a[2], out_b[2]) and torch.equal(out_a[0][0], out_b[0][0])
    assert float(out_a[4]["episode_steps"].max()) == 27


def test_fused_policy_kernel_matches_torch_fp32_reference(built, cuda_device):
    """sbr_policy_mlp (one launch on the SoA observation rows) against the plain torch fp32 expression of the same
    two-layer perceptron: fp32 rounding apart (sum order, tanh / exp implementations), identical actions."""
    n = 5000                                                     # ragged: not a multiple of the CTA size
    policy = rollout.TinyPolicy(cuda_device)
    g = torch.Generator(device=cuda_device).manual_seed(9)
    obs_do = torch.rand((9, n), dtype=torch.float64, device=cuda_device, generator=g) * 2 - 0.5
    obs_ec = torch.rand((9, n), dtype=torch.float64, device=cuda_device, generator=g) * 2 - 0.5
    ref = policy.forward_torch(obs_do, obs_ec)
    out = torch.full((2, n), float("nan"), dtype=torch.float64, device=cuda_device)
    got = policy.act_into(obs_do, obs_ec, out)
    assert got.data_ptr() == out.data_ptr() and bool(torch.isfinite(out).all())
    assert float((out - ref).abs().max()) <= 2e-5               # outputs span 6.5 / 13: ~1e-6 relative in fp32
    assert float(out[0].min()) >= 0.5 and float(out[0].max()) <= 7.0 and float(out[1].max()) <= 14.0
    assert torch.equal(policy.forward_soa(obs_do, obs_ec), out)
    # argument errors surface as exceptions, not as a silent fallback
    with pytest.raises(Exception):
        core.policy_mlp(obs_do, obs_ec, policy.w1t.double(), policy.w2t, policy.lo_flat, policy.span_flat, out)


@pytest.mark.parametrize("K", [8, 5])
def test_fused_rollout_matches_stepwise_rollout(built, cuda_device, K):
    """sbr_os_rollout_k (K steps per launch, policy evaluated in-kernel on the registers that hold the state) against
    the step-by-step rollout [sbr_policy_mlp, sbr_os_step]: returns, final state, per-step rewards, set-points and
    observations are bit-identical."""
    n = 300                                                      # ragged tiles, odd number of launches for K = 5
    policy = rollout.TinyPolicy(cuda_device)
    env_e = SbrOsVecEnv(n, device=cuda_device, seed=5, mode="dp45")
    eager = rollout.collect_episode(env_e, policy, store=True)
    env_f = SbrOsVecEnv(n, device=cuda_device, seed=5, mode="dp45")
    fused = rollout.collect_episode_fused(env_f, policy, K=K, store=True)
    assert bool(fused["all_done"]) and fused["steps"] == 463
    assert torch.equal(fused["returns"], eager["returns"])
    assert torch.equal(fused["rewards"], eager["rewards"])
    assert torch.equal(torch.nan_to_num(env_f.buf.st), torch.nan_to_num(env_e.buf.st))
    assert torch.equal(env_f.buf.obs_do, env_e.buf.obs_do) and torch.equal(env_f.buf.obs_ec, env_e.buf.obs_ec)
    # the logged set-points are what the stand-alone policy kernel computes from the logged observations
    acts, obs = fused["actions"], fused["observations"]
    chk = torch.empty((2, n), dtype=torch.float64, device=cuda_device)
    for k in (0, 17, 200, 461):
        policy.act_into(obs[k, :9].contiguous(), obs[k, 9:].contiguous(), chk)
        assert torch.equal(chk, acts[k + 1]), k
    assert int(fused["status"].max()) == 0
