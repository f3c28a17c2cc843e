"""Strong scaling of ONE frame (configs[1]: 5,242,880 rays) over the ranks of a torchrun launch, gather included.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      scripts/strong_frame.py

Every rank holds the full entry buffers (as a Blender process per GPU would after generating the camera rays),
integrates its interleaved shard and the exit buffers are gathered on rank 0 (distributed.trace_sharded).
Prints, for chunks in (1, 2, 4, 8), the frame latency (CUDA events on rank 0, max over ranks) so the overlap of
the gather with the integration can be read directly.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import api, distributed as D, raygen  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cam = api.make_camera(raygen.CFG_CAMERA_POS, raygen.look_at_rotation(raygen.CFG_CAMERA_POS), 1280, 1024,
                          raygen.CFG_FOV, raygen.CFG_FOV, seed=raygen.CFG_SEED, jitter="philox")
    pos, d, _ = api.generate_rays(cam, 4 * 1280 * 1024, raygen.CFG_R_SPHERE, device=local)   # device generator, same on all ranks
    kw = dict(M=raygen.CFG_M, r_sphere=raygen.CFG_R_SPHERE, rtol=1e-3, atol=1e-6)
    res = {}
    quick = os.environ.get("BHG_STRONG_QUICK", "0") == "1"   # fewer variants: an 8-GPU box is charged 8x
    for chunks in ((2,) if quick else (1, 2, 4, 8)):
        times = []
        for it in range(8):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = D.trace_sharded(pos, d, chunks=chunks, **kw)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it >= 3:
                times.append(float(t))
        res[chunks] = float(np.median(times))
    # peer-memory route: exit states stored straight into rank 0's HBM by every GPU's trace kernel
    frame = D.PeerFrame(pos.shape[0], owner=0)
    peer = {}
    variants = (("stores", 0, 1), ("stores", 1280, 1), ("copy", 0, 1), ("copy", 1280, 1), ("copy", 1280, 2),
                ("copy", 1280, 4), ("copy", 1280, 8))
    if quick:
        variants = (("stores", 0, 1), ("copy", 1280, 2))
    for route, width, chunks in variants:
        times = []
        for it in range(10):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pout = D.trace_sharded_peer(pos, d, frame, image_width=width, route=route, chunks=chunks, **kw)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it >= 3:
                times.append(float(t))
        peer[f"{route}/w{width}/c{chunks}"] = float(np.median(times))
    same = None
    if rank == 0:
        same = all(torch.equal(a, b) for a, b in zip(pout, out))
    # the same call captured in a CUDA graph (one launch per frame instead of ~25 host calls)
    graphs = {}
    if os.environ.get("BHG_STRONG_GRAPHS", "0") == "1":
        for route, width, chunks in (("stores", 0, 1), ("copy", 1280, 2)):
            try:
                if rank == 0:
                    pout[2].fill_(-9)
                g = D.PeerFrameGraph(pos, d, frame, image_width=width, route=route, chunks=chunks, **kw)
                times = []
                for it in range(10):
                    if world > 1:
                        dist.barrier()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    gout = g.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    if it >= 3:
                        times.append(float(t))
                graphs[f"{route}/w{width}/c{chunks}"] = float(np.median(times))
                if rank == 0:
                    graphs[f"{route}/w{width}/c{chunks}/equal"] = all(torch.equal(a, b) for a, b in zip(gout, out))
                g.close()
                del g, gout
            except Exception as e:  # report instead of losing the other numbers
                graphs[f"{route}/error"] = repr(e)[:300]
    frame.close()
    if rank == 0:
        n = pos.shape[0]
        st = out[2]
        print(json.dumps({"n_gpus": world, "rays": n, "gather_frame_ms_by_chunks": res, "peer_frame_ms": peer, "peer_frame_graph_ms": graphs,
                          "peer_equals_gather": same, "best_rays_per_s": n / (min(peer.values()) * 1e-3),
                          "status_counts": torch.bincount(st.to(torch.int64), minlength=6).tolist()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
