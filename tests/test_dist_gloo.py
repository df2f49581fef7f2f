"""Multi-process plumbing on CPU (gloo, world_size 2): env-index sharding, the reward-statistics gather (the
only collective of a rollout) and the per-env return gather with uneven shards."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    from gym_sbr2_b200 import dist
    r, w, _ = dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = dist.shard_range(n_total, rank, world)
    rewards = torch.arange(lo, hi, dtype=torch.float64) * 0.5 - 3.0       # a function of the GLOBAL env index
    local = torch.tensor([rewards.sum(), (rewards ** 2).sum(), rewards.min(), rewards.max(), float(hi - lo)],
                         dtype=torch.float64)
    stats = dist.gather_stats(local)
    combined = dist.combine_stats(stats)
    allr = dist.gather_rewards(rewards, n_total)
    torch.save(dict(stats=stats, combined=combined, allr=allr, lo=lo, hi=hi), os.path.join(out_dir, "r%d.pt" % rank))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 11])
def test_reward_gather_world2(tmp_path, n_total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    ref = np.arange(n_total) * 0.5 - 3.0
    shards = []
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), "r%d.pt" % r))
        shards.append((d["lo"], d["hi"]))
        assert np.array_equal(d["allr"].numpy(), ref)                       # same on every rank, global order
        c = d["combined"]
        assert c["count"] == n_total and c["min"] == ref.min() and c["max"] == ref.max()
        assert abs(c["mean"] - ref.mean()) < 1e-12 and abs(c["std"] - ref.std()) < 1e-12
        assert d["stats"].shape == (world, 5)
    assert shards[0][0] == 0 and shards[0][1] == shards[1][0] and shards[1][1] == n_total


def test_shard_range_partitions_exactly():
    from gym_sbr2_b200 import dist
    for n in (1, 7, 8, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 4, 8):
            blocks = [dist.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
