"""Shard invariance ("same per-env seeds => same per-env results, whatever the world size"): the
counter-based influent sampler, the SoA row permutation, and whole rollouts split over shards / ranks reproducing
the single-batch run bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from gym_sbr2_b200 import _abi, core, dist, influent, rollout
from gym_sbr2_b200.vec_env import SbrOsVecEnv, SbrV2VecEnv, SbrV4VecEnv
from oracle import philox_ref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_philox_normals_match_numpy_restatement(built, cuda_device):
    n, seed = 1500, 0x1234_5678_9ABC_DEF0
    for off, ep in ((0, 0), (10 ** 6 + 3, 5), (2 ** 33 + 1, 2)):            # 64-bit env indices reach the counter
        z = core.philox_normals(n, cuda_device, seed, env_offset=off, epoch0=ep).cpu().numpy()
        ref = philox_ref.normals(seed, off + np.arange(n), ep)
        assert np.allclose(z, ref, rtol=0, atol=2e-14)                     # libm vs CUDA log / sincospi: ulps
    assert abs(z.mean()) < 2e-2 and abs(z.std() - 1.0) < 2e-2


def test_influent_sample_is_mix_of_its_own_normals_bit_for_bit(built, cuda_device):
    """sbr_influent_sample == sbr_influent_mix(sbr_philox_normals): the mixing arithmetic is the one already pinned
    bit-exactly to the reference's numpy arithmetic; the sampler only changes where the 48 normals come from."""
    n, seed, off = 777, 99, 12345
    z = core.philox_normals(n, cuda_device, seed, env_offset=off, epoch0=3)
    for sw in range(8):
        got = core.influent_sample(n, cuda_device, seed, env_offset=off, scenario=sw, epoch0=3)
        assert torch.equal(got, core.influent_mix(sw, z)), sw
        ref = np.stack([influent.mix_numpy(sw, z[:, i].cpu().numpy()) for i in range(0, n, 97)], axis=1)
        assert np.array_equal(got[:, ::97].cpu().numpy(), ref)             # and with the reference's arithmetic
    # scenario = -1: drawn per env, uniform on 0..7, as the numpy restatement of the counter says
    scn = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    got = core.influent_sample(n, cuda_device, seed, env_offset=off, scenario=-1, epoch0=3, scenario_out=scn)
    assert np.array_equal(scn.cpu().numpy(), philox_ref.scenario(seed, off + np.arange(n), 3))
    for sw in range(8):
        m = scn == sw
        assert bool(m.any()) and torch.equal(got[:, m], core.influent_mix(sw, z)[:, m])


def test_sampler_is_invariant_to_sharding_and_honours_mask_and_epoch(built, cuda_device):
    n, seed = 1000, 7
    full = core.influent_sample(n, cuda_device, seed, scenario=6)
    parts = [core.influent_sample(hi - lo, cuda_device, seed, env_offset=lo, scenario=6)
             for lo, hi in ((0, 1), (1, 400), (400, 1000))]
    assert torch.equal(torch.cat(parts, dim=1), full)
    # per-env epoch: only masked envs draw, their epoch advances, everything else is left untouched
    epoch = torch.zeros(n, dtype=torch.int64, device=cuda_device)
    out = torch.full((14, n), -1.0, dtype=torch.float64, device=cuda_device)
    mask = (torch.arange(n, device=cuda_device) % 3 == 0).to(torch.uint8)
    core.influent_sample(n, cuda_device, seed, scenario=6, epoch=epoch, mask=mask, out=out)
    assert torch.equal(out[:, mask.bool()], full[:, mask.bool()]) and bool((out[:, ~mask.bool()] == -1.0).all())
    assert torch.equal(epoch, mask.to(torch.int64))
    core.influent_sample(n, cuda_device, seed, scenario=6, epoch=epoch, out=out)       # everybody draws again
    ep1 = core.influent_sample(n, cuda_device, seed, scenario=6, epoch0=1)
    assert torch.equal(out[:, mask.bool()], ep1[:, mask.bool()]) and torch.equal(out[:, ~mask.bool()], full[:, ~mask.bool()])
    assert torch.equal(epoch, mask.to(torch.int64) + 1)
    with pytest.raises(_abi.SbrLibraryError):
        core.influent_sample(n, cuda_device, seed, scenario=8)


def test_permute_rows_gather_and_scatter(built, cuda_device):
    n = 5000
    g = torch.Generator(device=cuda_device).manual_seed(1)
    perm = torch.randperm(n, device=cuda_device, generator=g)
    a = torch.rand((14, n), dtype=torch.float64, device=cuda_device, generator=g)
    big = torch.rand((3, n + 40), dtype=torch.float64, device=cuda_device, generator=g)
    b = big[:, :n]                                                          # ld > n
    c = torch.randint(0, 1 << 30, (n,), dtype=torch.int32, device=cuda_device, generator=g)
    d = torch.randint(0, 1 << 30, (2, n), dtype=torch.int32, device=cuda_device, generator=g)
    ga, gb, gc, gd = torch.empty_like(a), torch.empty((3, n), dtype=torch.float64, device=cuda_device), \
        torch.empty_like(c), torch.empty_like(d)
    core.permute_rows(perm, [(a, ga), (b, gb), (c, gc), (d, gd)])
    assert torch.equal(ga, a[:, perm]) and torch.equal(gb, b[:, perm]) and torch.equal(gc, c[perm]) \
        and torch.equal(gd, d[:, perm])
    sa, sc = torch.empty_like(a), torch.empty_like(c)
    core.permute_rows(perm, [(ga, sa), (gc, sc)], scatter=True)             # scatter undoes the gather
    assert torch.equal(sa, a) and torch.equal(sc, c)
    with pytest.raises(ValueError):
        core.permute_rows(perm, [(a, ga)] * 9)


def _os_returns(n, off, seed, device, steps=None):
    env = SbrOsVecEnv(n, device=device, seed=seed, mode="dp45", env_offset=off)
    ep = rollout.collect_episode(env, rollout.TinyPolicy(device), max_steps=steps)
    return ep["returns"], env


def test_shards_reproduce_the_single_batch_rollout_bit_for_bit(built, cuda_device):
    """BASELINE config 5 in small: one SBROS-v1 episode of 1536 envs driven by the policy, as one batch and as three
    uneven shards (warp-unaligned boundaries): per-env returns, final states and influent identical."""
    n, seed = 1536, 4242
    full, env = _os_returns(n, 0, seed, cuda_device)
    got, infl, st = [], [], []
    for lo, hi in ((0, 500), (500, 1037), (1037, n)):
        r, e = _os_returns(hi - lo, lo, seed, cuda_device)
        got.append(r); infl.append(e.influent); st.append(e.buf.st)
    assert torch.equal(torch.cat(infl, dim=1), env.influent)
    assert torch.equal(torch.cat(st, dim=1).view(torch.int64), env.buf.st.view(torch.int64))
    assert torch.equal(torch.cat(got), full)
    assert float(full.std()) > 0


def test_v2_and_v4_shards_reproduce_the_single_batch(built, cuda_device):
    n, seed = 1000, 17
    g = torch.Generator(device=cuda_device).manual_seed(5)
    act = torch.rand((n, 3), dtype=torch.float64, device=cuda_device, generator=g)
    for mode in ("rk4", "dp45"):
        env = SbrV2VecEnv(n, device=cuda_device, seed=seed, mode=mode)
        env.reset()
        _, reward, _, info = env.step(act)
        rw, xl = reward.clone(), info["x_last"].clone()
        parts = []
        for lo, hi in ((0, 333), (333, n)):
            e = SbrV2VecEnv(hi - lo, device=cuda_device, seed=seed, mode=mode, env_offset=lo)
            e.reset()
            _, r, _, i = e.step(act[lo:hi])
            parts.append((r.clone(), i["x_last"].clone()))
        assert torch.equal(torch.cat([p[0] for p in parts]), rw), mode
        assert torch.equal(torch.cat([p[1] for p in parts], dim=1), xl), mode
    a4 = torch.rand((n, 1), dtype=torch.float64, device=cuda_device, generator=g) * 0.1
    def run(lo, hi):
        e = SbrV4VecEnv(hi - lo, device=cuda_device, seed=seed, env_offset=lo)
        e.reset()
        for _ in range(40):
            obs, r, _, _ = e.step(a4[lo:hi])
        return obs.clone(), r.clone(), e.scenario.clone()
    full = run(0, n)
    halves = [run(0, 411), run(411, n)]
    for k in range(3):
        assert torch.equal(torch.cat([h[k] for h in halves]), full[k])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_worker(rank, world, port, n_total, seed, steps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    ngpu = torch.cuda.device_count()
    backend = "nccl" if ngpu >= world else "gloo"          # one GPU: both ranks share it, the gather goes over gloo
    device = torch.device("cuda", rank % ngpu)
    if backend == "nccl":
        dist.init_from_env(backend="nccl")
    else:
        torch.cuda.set_device(device)
        torch.distributed.init_process_group(backend="gloo", rank=rank, world_size=world)
    lo, hi = dist.shard_range(n_total, rank, world)
    returns, _ = _os_returns(hi - lo, lo, seed, device, steps=steps)
    allr = dist.gather_rewards(returns, n_total)
    if rank == 0:
        torch.save(dict(allr=allr.cpu(), backend=backend), os.path.join(out_dir, "allr.pt"))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_ranks_gather_the_single_rank_returns(built, cuda_device, tmp_path):
    """Two processes (one per GPU over NCCL when two GPUs are visible; otherwise both on cuda:0 with the gather over
    gloo), SbrOsVecEnv shards keyed by global env index + dist.gather_rewards == the single-process run."""
    n_total, seed, steps = 777, 31, 120
    full, _ = _os_returns(n_total, 0, seed, cuda_device, steps=steps)
    mp.spawn(_rank_worker, args=(2, _free_port(), n_total, seed, steps, str(tmp_path)), nprocs=2, join=True)
    d = torch.load(os.path.join(str(tmp_path), "allr.pt"))
    assert torch.equal(d["allr"], full.cpu()), d["backend"]
