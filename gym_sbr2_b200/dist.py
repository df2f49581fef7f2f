"""Multi-GPU plumbing: one process per GPU (torchrun), env instances sharded by contiguous index blocks, NO
collective on the step path.  NCCL (or gloo on CPU for tests) is used only to gather episode-reward statistics
and, on request, the per-env returns."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n_total, rank, world):
    """Contiguous env-index block [lo, hi) owned by `rank`; blocks differ by at most one env."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def combine_stats(stats):
    """stats: [world, 5] rows of [sum, sumsq, min, max, count] -> dict(mean, std, min, max, count)."""
    s = stats.to(torch.float64)
    cnt = float(s[:, 4].sum())
    tot, tot2 = float(s[:, 0].sum()), float(s[:, 1].sum())
    mean = tot / cnt if cnt else float("nan")
    var = max(tot2 / cnt - mean * mean, 0.0) if cnt else float("nan")
    return dict(mean=mean, std=var ** 0.5, min=float(s[:, 2].min()), max=float(s[:, 3].max()), count=cnt)


def gather_stats(local_stats, async_op=False):
    """all_gather of the per-rank 5-number reward statistics (the only collective of a rollout).  Returns the
    [world, 5] tensor (and the work handle when async_op)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        out = local_stats.reshape(1, 5).clone()
        return (out, None) if async_op else out
    out = torch.empty((world, 5), dtype=local_stats.dtype, device=local_stats.device)
    work = dist.all_gather_into_tensor(out, local_stats.reshape(1, 5).contiguous(), async_op=async_op)
    return (out, work) if async_op else out


def gather_rewards(local_rewards, n_total):
    """all_gather of per-env episode returns (variable shard sizes are padded to the largest shard)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local_rewards.clone()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    # gloo (CPU tests, or two ranks sharing one GPU) gathers host tensors; NCCL gathers in place on the device
    dev = local_rewards.device if dist.get_backend() == "nccl" else torch.device("cpu")
    pad = torch.full((width,), float("nan"), dtype=local_rewards.dtype, device=dev)
    pad[: local_rewards.shape[0]] = local_rewards
    out = [torch.empty((width,), dtype=local_rewards.dtype, device=dev) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([out[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)]).to(local_rewards.device)
