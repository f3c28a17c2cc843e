"""CPU tests of the multi-GPU host logic: world_size-2 gloo process group, interleaved ray sharding and the
gather of exit buffers to the frame-owning rank.  The tracer is injected (the oracle's C port) because the
product tracer needs a GPU; the partition / gather code under test is the product's."""
import os
import socket

import numpy as np
import pytest

from blackhole_geodesic_calculator_b200 import distributed as D


def test_partitions():
    n, w = 1003, 4
    seen = np.concatenate([D.interleaved_indices(n, r, w) for r in range(w)])
    assert sorted(seen.tolist()) == list(range(n))
    assert [D.shard_size(n, r, w) for r in range(w)] == [251, 251, 251, 250]
    assert D.frames_for_rank(10, 1, 4) == [1, 5, 9]
    assert sum(len(D.frames_for_rank(100, r, 8)) for r in range(8)) == 100


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from blackhole_geodesic_calculator_b200 import raygen
        from oracle import port as oracle_port

        def tracer(pos, d, **kw):
            o = oracle_port.trace(pos, d, nthreads=1, **kw)
            return o["exit_pos"], o["exit_dir"], o["status"]

        pos, d = raygen.config_bundle(32, 32, 1)
        pos, d = pos[:n], d[:n]
        out = D.trace_sharded(pos, d, dst=0, tracer=tracer, chunks=3, rtol=1e-3, atol=1e-6)
        if rank == 0:
            ref = tracer(pos, d, rtol=1e-3, atol=1e-6)
            ok = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(out, ref))
            q.put(("ok" if ok else "mismatch", out[2].shape[0]))
        else:
            q.put(("none" if out is None else "unexpected", 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1024, 1001])  # even and ragged shards
def test_trace_sharded_world2_gloo(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [("none", 0), ("ok", n)]


@pytest.mark.parametrize("n,world,width", [(8 * 48 * 5, 2, 48), (1000, 3, 0), (8 * 48 * 5, 8, 48), (31, 4, 0)])
def test_shard_order_partitions_the_frame(n, world, width):
    """Every ray belongs to exactly one rank; shards differ by at most one 32-slot group."""
    parts = [D.shard_order(n, r, world, width) for r in range(world)]
    assert all(p.dtype == np.int32 for p in parts)
    assert sorted(np.concatenate(parts).tolist()) == list(range(n))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 32
    if width:   # a group of 32 slots is a 4 x 8 pixel tile
        y, x = np.divmod(parts[0][:32].astype(np.int64), width)
        assert x.max() - x.min() == 3 and y.max() - y.min() == 7


@pytest.mark.parametrize("n,world,width", [(8 * 48 * 5, 2, 48), (20000, 3, 0), (8 * 48 * 5, 8, 48), (31, 4, 0),
                                           (8192 * 3 + 5, 2, 0)])
def test_band_plan_partitions_the_frame(n, world, width):
    """Copy route: whole bands per rank, every ray exactly once, tile hint only when every band is 8 full rows."""
    seen = []
    for r in range(world):
        band, mine, m, tiles_ok = D.band_plan(n, r, world, width)
        idx = (mine[:, None] * band + np.arange(band)[None, :]).reshape(-1)
        idx = idx[idx < n]
        assert idx.size == m and (mine % world == r).all()
        assert tiles_ok == bool(width) and (not tiles_ok or band == 8 * width)
        seen += idx.tolist()
    assert sorted(seen) == list(range(n))


def _subgroup_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from blackhole_geodesic_calculator_b200 import raygen
        from oracle import port as oracle_port

        def tracer(pos, d, **kw):
            o = oracle_port.trace(pos, d, nthreads=1, **kw)
            return o["exit_pos"], o["exit_dir"], o["status"]

        group = dist.new_group([1, 2])       # every rank must take part in creating it
        if rank == 0:
            q.put(("outside", 0))
            return
        pos, d = raygen.config_bundle(32, 32, 1)
        pos, d = pos[:301], d[:301]
        # dst is a rank OF THE GROUP: group rank 1 is global rank 2
        out = D.trace_sharded(pos, d, group=group, dst=1, tracer=tracer, rtol=1e-3, atol=1e-6)
        if rank == 2:
            ref = tracer(pos, d, rtol=1e-3, atol=1e-6)
            ok = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(out, ref))
            q.put(("ok" if ok else "mismatch", 2))
        else:
            q.put(("none" if out is None else "unexpected", 1))
    finally:
        dist.destroy_process_group()


def test_trace_sharded_subgroup_dst_is_a_group_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_subgroup_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [("none", 1), ("ok", 2), ("outside", 0)]
