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


def trace_sharded(entry_pos, entry_dir, *, group=None, dst=0, tracer=None, chunks=None, **trace_kw):
    """Trace one frame's rays across all ranks of `group` and gather the exit buffers on rank `dst`.

    Every rank passes the same full `entry_pos` / `entry_dir` ([N,3]; numpy arrays, or torch CUDA tensors for
    the NCCL path).  Each rank integrates its interleaved shard; rank `dst` returns
    (exit_pos[N,3], exit_dir[N,3], status[N]) in the original ray order, every other rank returns None.
    The shard is processed in `chunks` pieces (default 4 on CUDA tensors, 1 otherwise): the gather of piece c is
    issued asynchronously and overlaps the integration of piece c + 1, so at 8 GPUs the 52 B/ray that converge on
    the frame owner hide behind the compute instead of following it.
    `tracer` defaults to the CUDA `api.trace`; the CPU (gloo) tests inject the oracle.
    """
    import torch
    import torch.distributed as dist

    tracer = tracer or _default_tracer
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    is_torch = isinstance(entry_pos, torch.Tensor)
    n = entry_pos.shape[0]
    if is_torch:
        idx = torch.arange(rank, n, world, device=entry_pos.device)
        pos, d = entry_pos.index_select(0, idx).contiguous(), entry_dir.index_select(0, idx).contiguous()
    else:
        idx = interleaved_indices(n, rank, world)
        pos, d = np.ascontiguousarray(entry_pos[idx]), np.ascontiguousarray(entry_dir[idx])
    if world == 1:
        return tuple(tracer(pos, d, **trace_kw)[:3])
    m_max = (n + world - 1) // world          # largest shard; smaller shards are padded
    if chunks is None:
        chunks = 4 if is_torch else 1
    chunks = max(1, min(int(chunks), m_max))
    piece = (m_max + chunks - 1) // chunks
    dev = entry_pos.device if is_torch else torch.device("cpu")
    m = pos.shape[0]
    works, gathered = [], []
    for c in range(chunks):
        lo, hi = c * piece, min((c + 1) * piece, m_max)
        if lo >= hi:
            break
        out6 = torch.full((hi - lo, 6), float("nan"), dtype=torch.float64, device=dev)
        sts = torch.full((hi - lo,), -1, dtype=torch.int32, device=dev)
        a, b = min(lo, m), min(hi, m)
        if b > a:
            ep, ed, st = tracer(pos[a:b], d[a:b], **trace_kw)[:3]
            if not is_torch:
                ep, ed, st = torch.from_numpy(ep), torch.from_numpy(ed), torch.from_numpy(st)
            out6[:b - a, :3], out6[:b - a, 3:], sts[:b - a] = ep, ed, st
        if rank == dst:
            g6 = [torch.empty_like(out6) for _ in range(world)]
            gs = [torch.empty_like(sts) for _ in range(world)]
        else:
            g6 = gs = None
        works.append(dist.gather(out6, g6, dst=dst, group=group, async_op=True))
        works.append(dist.gather(sts, gs, dst=dst, group=group, async_op=True))
        gathered.append((g6, gs, out6, sts))  # keep the send buffers alive until the gathers complete
    for w in works:
        w.wait()
    if rank != dst:
        return None
    # un-interleave: row j of rank r's shard is ray r + j * world
    full6 = torch.cat([torch.stack(g6, dim=1) for g6, _, _, _ in gathered], dim=0).reshape(-1, 6)[:n]
    fulls = torch.cat([torch.stack(gs, dim=1) for _, gs, _, _ in gathered], dim=0).reshape(-1)[:n]
    exit_pos, exit_dir = full6[:, :3].contiguous(), full6[:, 3:].contiguous()
    if is_torch:
        return exit_pos, exit_dir, fulls
    return exit_pos.numpy(), exit_dir.numpy(), fulls.numpy()


# ---------------------------------------------------------------------------------------------------------------
# Peer-memory frame: every GPU stores its exit states straight into the frame owner's HBM over NVLink while it
# integrates (CUDA IPC mapping, include/bhgeo.h bhg_ipc_*), so the gather disappears into the trace kernel.
# ---------------------------------------------------------------------------------------------------------------

def shard_order(n: int, rank: int, world: int, image_width: int = 0) -> np.ndarray:
    """int32 ray indices integrated by `rank`, in queue order.

    Groups of 32 consecutive queue slots (what one warp fetches when it starts) are dealt round-robin to the ranks,
    so every GPU sees the same mix of cheap and near-shadow rays; with `image_width` (row-major image, width a
    multiple of 4, n a multiple of 8 rows) a group is one 4 x 8 pixel tile, the same coherence mapping the
    single-GPU kernel applies (csrc/trace_kernel.cuh slot_to_ray)."""
    if n >= 2 ** 31:
        raise ValueError("shard_order: n must fit int32")
    slots = np.arange(n, dtype=np.int64)
    mine = slots[(slots >> 5) % world == rank]
    if image_width > 0 and image_width % 4 == 0 and n % (8 * image_width) == 0:
        band_sz = 8 * image_width
        band, t = np.divmod(mine, band_sz)
        tile, lane = t >> 5, t & 31
        mine = band * band_sz + (lane >> 2) * image_width + (tile << 2) + (lane & 3)
    return mine.astype(np.int32)


class _DevicePointerView:
    """Minimal __cuda_array_interface__ carrier so torch can wrap library-owned device memory without a copy."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerFrame:
    """Exit buffers (exit_pos[n,3] f64, exit_dir[n,3] f64, status[n] i32) of one frame, resident in the HBM of rank
    `owner` and mapped into every other rank of `group` (one process per GPU).  Collective: construct and close it
    on all ranks."""

    def __init__(self, n: int, *, group=None, owner: int = 0, device=None):
        import ctypes

        import torch
        import torch.distributed as dist
        from . import _lib

        self._lib = _lib.load()
        self._check = _lib.check
        self.n, self.group, self.owner = int(n), group, int(owner)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.cuda.current_device() if device is None else int(device)
        align = lambda b: (b + 255) & ~255
        self._off_dir = align(self.n * 24)
        self._off_status = self._off_dir + align(self.n * 24)
        total = self._off_status + self.n * 4
        base = ctypes.c_void_p()
        handle = None
        self.is_owner = self.rank == self.owner
        if self.is_owner:
            self._check(self._lib.bhg_device_alloc(total, self.device, ctypes.byref(base)))
            if self.world > 1:
                buf = ctypes.create_string_buffer(64)
                self._check(self._lib.bhg_ipc_export(base, self.device, buf))
                handle = buf.raw
        if self.world > 1:
            box = [handle]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, self.owner) if group else self.owner,
                                       group=group)
            if not self.is_owner:
                self._check(self._lib.bhg_ipc_open(box[0], self.device, ctypes.byref(base)))
        self._base = base.value
        self._order_cache = {}
        self._token = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", self.device))

    pos_ptr = property(lambda self: self._base)
    dir_ptr = property(lambda self: self._base + self._off_dir)
    status_ptr = property(lambda self: self._base + self._off_status)

    def tensors(self):
        """(exit_pos, exit_dir, status) torch views of the frame; owner only."""
        import torch

        if not self.is_owner:
            return None
        dev = torch.device("cuda", self.device)
        wrap = lambda p, shape, ts: torch.as_tensor(_DevicePointerView(p, shape, ts), device=dev)
        return (wrap(self.pos_ptr, (self.n, 3), "<f8"), wrap(self.dir_ptr, (self.n, 3), "<f8"),
                wrap(self.status_ptr, (self.n,), "<i4"))

    def order(self, image_width=0):
        import torch

        key = int(image_width)
        if key not in self._order_cache:
            o = shard_order(self.n, self.rank, self.world, key)
            self._order_cache[key] = torch.from_numpy(o).to(torch.device("cuda", self.device))
        return self._order_cache[key]

    def fence(self):
        """Stream-ordered all-ranks fence (1-element all-reduce; no host synchronisation)."""
        import torch.distributed as dist

        if self.world > 1:
            dist.all_reduce(self._token, group=self.group)

    def close(self):
        import torch

        if self._base is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)  # nobody still writes into the mapping
        if self.is_owner:
            self._check(self._lib.bhg_device_free(self._base, self.device))
        else:
            self._check(self._lib.bhg_ipc_close(self._base, self.device))
        self._base = None


def trace_sharded_peer(entry_pos, entry_dir, frame: PeerFrame, *, image_width=0, fence_before=True, **trace_kw):
    """Trace one frame across all ranks with the exit states written directly into `frame` (the owner's HBM).

    entry_pos / entry_dir: the frame's full [n,3] float64 CUDA tensors, present on every rank (each rank generates
    them from the camera).  Each rank integrates the rays of `frame.order(image_width)` and stores their exit
    states at their final position in the owner's buffers through the NVLink mapping - the kernel is the same
    bhg_trace_schwarzschild_f64 with remote out pointers and an `order` array.  A closing 1-element all-reduce
    orders the owner's stream after every rank's kernel; `fence_before` adds the same fence ahead of the kernel so
    that the previous frame's consumer on the owner has finished with the buffers before anyone overwrites them.
    Returns the owner's (exit_pos, exit_dir, status) views, None elsewhere.  Asynchronous on the current stream."""
    import torch
    from . import api

    if entry_pos.shape[0] != frame.n:
        raise ValueError("trace_sharded_peer: frame was built for a different ray count")
    dev = entry_pos.device
    order = frame.order(image_width)
    params = api.make_params(**trace_kw)
    if fence_before:
        frame.fence()
    if order.numel():
        api.trace_device(entry_pos.data_ptr(), entry_dir.data_ptr(), frame.pos_ptr, frame.dir_ptr, frame.status_ptr,
                         None, order.data_ptr(), order.numel(), api.LAYOUT_AOS, params, device=dev.index,
                         stream=torch.cuda.current_stream(dev).cuda_stream)
    frame.fence()
    return frame.tensors()
