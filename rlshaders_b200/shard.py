"""Multi-GPU partitioning of the path for the one-process-per-GPU harnesses (bench.py under torchrun).

Samples are independent, so configs 1-4 shard by contiguous index ranges with no data-path collective; the albedo sweep
shards the samples-per-pixel range of EVERY cell so the load is even, and the per-rank partial tables are summed with
one all-reduce.  The partition itself is the library's (`rls_multi_partition`, rlshaders_b200/csrc/rls_multi.cu -- the
function the single-process multi-device driver uses), so a torchrun job and `rls_driver --gpus N` shard identically;
it needs the shared library but no device.
"""
import ctypes as C

from . import _lib


def shard_range(total, rank, world):
    """Contiguous, near-equal slice [begin, end) of `total` items for `rank` of `world` (rls_multi_partition)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    b, e = C.c_uint64(), C.c_uint64()
    if _lib.load().rls_multi_partition(int(total), int(world), int(rank), C.byref(b), C.byref(e)) != 0:
        raise ValueError("rls_multi_partition rejected the arguments")
    return b.value, e.value


def spp_range(spp, rank, world):
    """The sweep's per-cell sample range owned by `rank` (same contract as shard_range)."""
    return shard_range(spp, rank, world)


def reduce_table(table, dist=None):
    """Sum the per-rank partial sweep tables in place (NCCL on GPU tensors, gloo on CPU tensors).  `dist` =
    torch.distributed when a process group is initialised, else None.  (The single-process form of the same step is
    rls_multi_albedo_sweep: kernel + ncclAllReduce in one CUDA graph per device.)"""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(table, op=dist.ReduceOp.SUM)
    return table
