"""Worker for tests/test_dist_cpu.py: one rank of a world_size-2 gloo job (CPU).  Exercises the only
multi-GPU logic the path has: contiguous sharding of the global env range and the all-reduce of the five
episode-return statistics."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_supplychain_b200.dist import allreduce_return_stats, shard_range, summarize_stats  # noqa: E402


def main():
    out_dir = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_total = 1001                                    # not divisible by the world size on purpose
    first, n_local = shard_range(n_total, rank, world)
    # every rank derives the same synthetic "episode returns" for the global env index, takes its shard
    returns = -1e6 * (1.0 + np.sin(np.arange(n_total) * 0.37))
    mine = returns[first:first + n_local]
    local = torch.tensor([len(mine), mine.sum(), (mine ** 2).sum(), mine.min(), mine.max()], dtype=torch.float64)
    glob = allreduce_return_stats(local)
    # bench.py's overlapped exchange: the all-gather of step k is consumed after step k+1 was submitted; the folded
    # totals must equal the plain all-reduce of the summed per-step records
    import bench

    class Ctx(object):
        pass
    ctx = Ctx()
    ctx.torch, ctx.dist, ctx.world, ctx.rank, ctx.dev = torch, dist, world, rank, torch.device("cpu")
    xch = bench.StatsExchange(ctx)
    chunks = np.array_split(mine, 5)
    for c in chunks:
        rec = torch.tensor([len(c), c.sum(), (c ** 2).sum(), c.min(), c.max()], dtype=torch.float64)
        xch.consume(xch.submit(rec))                  # start this step's exchange, fold the previous one
    pending_before_drain = xch.pending is not None
    xch.drain()
    res = {"rank": rank, "world": world, "first": first, "n_local": n_local, "global": glob.tolist(),
           "summary": list(summarize_stats(glob)), "overlapped": xch.total.tolist(), "pending_before_drain": pending_before_drain}
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as f:
        json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
