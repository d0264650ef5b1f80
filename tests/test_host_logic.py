"""CPU: host-side logic -- the synthetic generators' integer hash, index-range sharding,
and the N > 1 sweep reduction over a world_size-2 gloo group (the oracle stands in for the
kernel so the partition + all-reduce plumbing is what is under test)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from rlshaders_b200 import _abi as abi
from rlshaders_b200 import shard


def test_hash_uniform_numpy_matches_c():
    orc = ol.load_port()
    for seed, stream, first, lo, hi in ((0x5EED0001, 0, 0, 0.0, 1.0), (7, 3, 123456789012, 0.05, 2.5),
                                        (2**63 + 5, 50, 2**40, -1.0, 1.0)):
        a = orc.synth_uniform(4096, seed, stream, first, lo, hi)
        b = ol.hash_uniform(4096, seed, stream, first, lo, hi)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    u = ol.hash_uniform(1 << 20, 1, 0)
    assert u.min() >= 2.0 ** -24 and u.max() <= 1 - 2.0 ** -24     # never 0 or 1
    assert abs(u.mean() - 0.5) < 2e-3


def test_hash_is_index_addressed():
    """Slices are reproducible independent of how the range is partitioned."""
    full = ol.hash_uniform(10000, 42, 5)
    for world in (2, 4, 8):
        parts = [ol.hash_uniform(e - b, 42, 5, first_index=b) for b, e in
                 (shard.shard_range(10000, r, world) for r in range(world))]
        assert np.array_equal(np.concatenate(parts), full)


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 4096, 2**26 + 3):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard.shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _sweep_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = ol.load_port()
    orc.set_threads(1)
    grid = abi.SweepGrid(3, 4, 2, 0.05, 1.0, 1.0, 2.0)
    k0, k1 = shard.spp_range(96, rank, world)
    table = torch.from_numpy(orc.albedo_sweep(grid, 99, k0, k1))
    shard.reduce_table(table, dist)
    if rank == 0:
        ret.put(table.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_sweep_shards_sum_to_the_full_table_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sweep_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    reduced = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    orc = ol.load_port()
    grid = abi.SweepGrid(3, 4, 2, 0.05, 1.0, 1.0, 2.0)
    full = orc.albedo_sweep(grid, 99, 0, 96)
    # counts are exact; FP64 sums agree to rounding regardless of the partition
    assert np.array_equal(reduced[:, 3:], full[:, 3:])
    assert np.allclose(reduced, full, rtol=1e-12, atol=0)


def test_reduce_table_is_identity_without_a_group():
    t = torch.arange(10, dtype=torch.float64)
    assert torch.equal(shard.reduce_table(t.clone(), None), t)
    assert torch.equal(shard.reduce_table(t.clone(), dist), t)


# ------------------------------------------------------------------ argument validation of the Python mirror
def test_api_rejects_wrong_lengths_dtypes_and_placement():
    """Every pointer the mirror hands to the C ABI is checked first: a wrong length or a tensor in the wrong memory would
    be an out-of-bounds read / illegal address on the device (a dead CUDA context), not an exception."""
    import pytest
    import torch
    from rlshaders_b200 import api

    n = 8
    f = lambda *s: torch.zeros(*s, dtype=torch.float32)   # noqa: E731
    with pytest.raises(ValueError):
        api.ShadingBatch(f(3, n), f(3, n), f(3, n + 1), f(3, n))                       # N has another sample count
    with pytest.raises(ValueError):
        api.ShadingBatch(f(3, n), f(3, n), f(3, n), f(3, n), torch.zeros(n, dtype=torch.int32))   # backfacing must be uint8
    with pytest.raises(ValueError):
        api.ShadingBatch(f(3, n), f(3, n), f(3, n), f(3, n), torch.zeros(n + 1, dtype=torch.uint8))
    with pytest.raises(ValueError):
        api._f32rows(f(3, n), "wi", n + 1)
    with pytest.raises(ValueError):
        api._f32(f(n).double(), "rx", n)
    with pytest.raises(ValueError):
        api._i32(torch.zeros(n, dtype=torch.int64), "flags", n)

    class FakeCtx:                               # placement rules need only the context's device
        device = torch.device("cuda", 0)
    with pytest.raises(ValueError):
        api._f32(f(n), "rx", n, FakeCtx, host=False)                                   # CPU tensor to a device entry point
    api._f32(f(n), "rx", n, FakeCtx, host=True)                                        # ... fine for a *_host form
    sg = api.ShadingBatch(f(3, n), f(3, n), f(3, n), f(3, n))
    with pytest.raises(ValueError):
        sg.placed(FakeCtx, host=False)
    with pytest.raises(ValueError):
        api._params_placed(dict(ior=f(n + 1)), n, FakeCtx, True)                       # per-sample parameter of the wrong length


def test_partition_is_the_librarys_and_survives_huge_totals():
    """rls_multi_partition (rlshaders_b200/csrc/rls_multi.cu; shard.shard_range delegates to it): exact partition also where
    total * k would overflow 64 bits."""
    total = (1 << 64) - 3
    ranges = [shard.shard_range(total, r, 7) for r in range(7)]
    assert ranges[0][0] == 0 and ranges[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert max(e - b for b, e in ranges) - min(e - b for b, e in ranges) <= 1
    assert [shard.shard_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]     # = r * 10 // 4


def test_compact_frames_numpy_restatement_round_trips():
    """oracle_lib.frame_from_quaternion restates the header's DEFINITION of the frame of a unit quaternion
    (include/rls_b200.h rls_shading_quat_soa); quaternion_from_frame is how a host would encode its frames.  Encoding
    random orthonormal frames and decoding them returns the frames to one rounding, orthonormal to 1e-6."""
    import numpy as np
    import oracle_lib as ol
    sg = ol.make_shading(50000, 0xF7A3E)
    q = ol.quaternion_from_frame(sg)
    assert q.dtype == np.float32 and np.abs(np.sum(q.astype(np.float64) ** 2, axis=0) - 1.0).max() < 2e-7
    d = ol.frame_from_quaternion(q)
    assert max(np.abs(d[k] - sg[k]).max() for k in d) < 1e-6
    U, V, N = (np.stack([d[a + c] for c in "xyz"]).astype(np.float64) for a in "UVN")
    for a, b in ((U, V), (U, N), (V, N)):
        assert np.abs((a * b).sum(0)).max() < 1e-6
    for a in (U, V, N):
        assert np.abs((a * a).sum(0) - 1.0).max() < 1e-6
    assert np.abs(np.cross(U.T, V.T).T - N).max() < 1e-6          # right-handed
    # the identity quaternion is the identity frame, exactly
    e = ol.frame_from_quaternion(np.array([[0.0], [0.0], [0.0], [1.0]], np.float32))
    assert [float(e[k][0]) for k in ("Ux", "Uy", "Uz", "Vx", "Vy", "Vz", "Nx", "Ny", "Nz")] == [1, 0, 0, 0, 1, 0, 0, 0, 1]
