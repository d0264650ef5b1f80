"""Multi-GPU partitioning of the path.

Samples are independent, so configs 1-4 shard by contiguous index ranges with no
data-path collective: rank r of W owns samples [r*n/W, (r+1)*n/W) of the index-addressed
synthetic stream (or, for weak scaling, its own n-sample slice starting at r*n).  The
albedo sweep shards the samples-per-pixel range of EVERY cell so the load is even, and
the per-rank partial tables are summed with one all-reduce.
"""


def shard_range(total, rank, world):
    """Contiguous, near-equal slice [begin, end) of `total` items for `rank` of `world`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return rank * total // world, (rank + 1) * total // world


def spp_range(spp, rank, world):
    """The sweep's per-cell sample range owned by `rank` (same contract as shard_range)."""
    return shard_range(spp, rank, world)


def reduce_table(table, dist=None):
    """Sum the per-rank partial sweep tables in place (NCCL on GPU tensors, gloo on CPU
    tensors).  `dist` = torch.distributed when a process group is initialised, else None."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(table, op=dist.ReduceOp.SUM)
    return table
