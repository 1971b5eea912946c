"""Multi-GPU plumbing: envs are independent, so the path shards with NO data-path collective.

One process per GPU (torchrun); rank r simulates the contiguous range of global env indices
[first_env, first_env + n_local).  Philox streams are keyed by the GLOBAL env index, so a trajectory does
not depend on how many GPUs the job uses.  The only collective is the reduction of the five
episode-return statistics (count, sum, sumsq, min, max) -- 40 bytes, NCCL over NVLink on GPUs, gloo in
the CPU tests.
"""
import math


def shard_range(n_total, rank, world_size):
    """Contiguous, balanced partition of range(n_total): returns (first_env, n_local)."""
    base, rem = divmod(int(n_total), int(world_size))
    n_local = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n_local


def allreduce_return_stats(stats, group=None):
    """stats: float64[5] tensor = (count, sum, sumsq, min, max) of one shard.  Returns the job-wide values on every
    rank.  ONE collective: the 40-byte records are all-gathered and reduced locally (SUM on the first three, MIN / MAX
    on the last two) -- pure latency, so one NCCL call instead of one per reduction operator."""
    import torch
    import torch.distributed as dist
    out = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        parts = [torch.empty_like(out) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, out, group=group)
        allr = torch.stack(parts)                      # [world, 5]
        out[0:3] = allr[:, 0:3].sum(0)
        out[3] = allr[:, 3].min()
        out[4] = allr[:, 4].max()
    return out


def summarize_stats(stats):
    """(count, mean, std, min, max) from the 5 raw statistics."""
    cnt, s, sq, mn, mx = [float(x) for x in stats.tolist()]
    if cnt == 0:
        return 0, float("nan"), float("nan"), float("nan"), float("nan")
    mean = s / cnt
    var = max(sq / cnt - mean * mean, 0.0)
    return int(cnt), mean, math.sqrt(var), mn, mx
