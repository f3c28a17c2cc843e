"""Multi-GPU distribution of the geodesic batch: one process per GPU (torch.distributed, NCCL over
NVLink / NVSwitch on the B200 box, gloo in the CPU tests).

Rays are independent (each depends only on its own entry state and the global scalars), so there is no
exchange during integration.  The reference has no counterpart (single process; its only parallelism is a
commented-out per-row `mp.Pool`, raytracer/RelativisticRenderEngine.py:210-216, and an offline multi-process
camera pre-run, raytracer/RelativisticRenderEngineCamEdition.py:216).

Two partitions:
  * within a frame: INTERLEAVED rays (ray i -> rank i mod G), so the expensive near-shadow rays (step counts
    vary 20x inside a frame) are spread over all GPUs instead of landing in one contiguous tile;
  * across an animation: by frame (frame f -> rank f mod G).
The only collective is the gather of the exit buffers (6 x f64 + 1 x i32 = 52 B/ray) to the rank that owns
the Blender frame.
"""
from __future__ import annotations

import numpy as np


def interleaved_indices(n: int, rank: int, world: int) -> np.ndarray:
    """Ray indices of `rank` under the interleaved partition."""
    return np.arange(rank, n, world, dtype=np.int64)


def shard_size(n: int, rank: int, world: int) -> int:
    return (n - rank + world - 1) // world if rank < n else 0


def frames_for_rank(n_frames: int, rank: int, world: int):
    """Animation frames integrated by `rank` (config 4: shard by frame)."""
    return list(range(rank, n_frames, world))


def _default_tracer(pos, d, **kw):
    from . import api
    return api.trace(pos, d, **kw)


def trace_sharded(entry_pos, entry_dir, *, group=None, dst=0, tracer=None, **trace_kw):
    """Trace one frame's rays across all ranks of `group` and gather the exit buffers on rank `dst`.

    Every rank passes the same full `entry_pos` / `entry_dir` ([N,3]; numpy arrays, or torch CUDA tensors for
    the NCCL path).  Each rank integrates its interleaved shard; rank `dst` returns
    (exit_pos[N,3], exit_dir[N,3], status[N]) in the original ray order, every other rank returns None.
    `tracer` defaults to the CUDA `api.trace`; the CPU (gloo) tests inject the oracle.
    """
    import torch
    import torch.distributed as dist

    tracer = tracer or _default_tracer
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    is_torch = isinstance(entry_pos, torch.Tensor)
    n = entry_pos.shape[0]
    m_max = (n + world - 1) // world
    if is_torch:
        idx = torch.arange(rank, n, world, device=entry_pos.device)
        pos, d = entry_pos.index_select(0, idx).contiguous(), entry_dir.index_select(0, idx).contiguous()
    else:
        idx = interleaved_indices(n, rank, world)
        pos, d = np.ascontiguousarray(entry_pos[idx]), np.ascontiguousarray(entry_dir[idx])
    ep, ed, st = tracer(pos, d, **trace_kw)[:3]
    if world == 1:
        return ep, ed, st
    # pack the shard into fixed-size buffers (pad to the largest shard) and gather on dst
    if is_torch:
        dev = entry_pos.device
        out6 = torch.full((m_max, 6), float("nan"), dtype=torch.float64, device=dev)
        sts = torch.full((m_max,), -1, dtype=torch.int32, device=dev)
        m = ep.shape[0]
        out6[:m, :3], out6[:m, 3:], sts[:m] = ep, ed, st
    else:
        dev = torch.device("cpu")
        out6 = torch.full((m_max, 6), float("nan"), dtype=torch.float64)
        sts = torch.full((m_max,), -1, dtype=torch.int32)
        m = ep.shape[0]
        out6[:m, :3], out6[:m, 3:], sts[:m] = torch.from_numpy(ep), torch.from_numpy(ed), torch.from_numpy(st)
    if rank == dst:
        g6 = [torch.empty_like(out6) for _ in range(world)]
        gs = [torch.empty_like(sts) for _ in range(world)]
    else:
        g6 = gs = None
    dist.gather(out6, g6, dst=dst, group=group)
    dist.gather(sts, gs, dst=dst, group=group)
    if rank != dst:
        return None
    # un-interleave: row j of rank r's shard is ray r + j * world
    full6 = torch.stack(g6, dim=1).reshape(m_max * world, 6)[:n]
    fulls = torch.stack(gs, dim=1).reshape(m_max * world)[:n]
    exit_pos, exit_dir = full6[:, :3].contiguous(), full6[:, 3:].contiguous()
    if is_torch:
        return exit_pos, exit_dir, fulls
    return exit_pos.numpy(), exit_dir.numpy(), fulls.numpy()
