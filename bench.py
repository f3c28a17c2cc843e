#!/usr/bin/env python
"""Benchmark of the hot path: batched Schwarzschild null-geodesic integration (BASELINE.json metric
"geodesic rays/s (1024^2 x 5 spp Schwarzschild)").

    python bench.py --gpus N --steps K --warmup W              # this framework (sm_100a kernel)
    python bench.py --impl reference --steps K --warmup W      # the reference's CPU method on host cores

One "step" = one pass of the hot path over one frame's batch of rays (config 2: 5 242 880 rays).  At N > 1
(torchrun) every rank integrates its own frame of the orbiting-camera animation (config 4: frames shard
across GPUs, no data-path collective) => weak scaling; value = rays of all ranks / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, SPP = 1024, 1024, 5
FLOP_FIXED, FLOP_PER_ATTEMPT = 230.0, 760.0  # FLOP(ray) = 230 + 760 * n_attempt
# the optional plane mode integrates 6 variables instead of 8: RHS 22 flops (csrc/geodesic_core.cuh Rhs<3>) instead of
# 40, Runge-Kutta algebra 6/8 of 520 -> 6 * 22 + 390 = 522 per attempt; fixed part without the theta conversions
FLOP_FIXED_PLANE, FLOP_PER_ATTEMPT_PLANE = 200.0, 522.0
NOMINAL_FP64_TFLOPS = 37.2                   # 148 SM x 64 FMA/clk x 2 x 1.965 GHz


def frame_camera_pos(frame_index: int):
    """Camera position of animation frame `frame_index` (config 4: azimuth += 3.6 deg per frame; frame 0 = config 2)."""
    from blackhole_geodesic_calculator_b200 import raygen
    az = math.radians(3.6 * frame_index)
    c0 = np.array(raygen.CFG_CAMERA_POS)
    ca, sa = math.cos(az), math.sin(az)
    return (ca * c0[0] - sa * c0[1], sa * c0[0] + ca * c0[1], c0[2])


def frame_rays(frame_index: int, width=W, height=H, spp=SPP, first=0, count=None):
    """Entry positions/directions of one frame of the orbiting-camera animation (host generator)."""
    from blackhole_geodesic_calculator_b200 import raygen
    return raygen.config_bundle(width, height, spp, jitter="philox", cam_pos=frame_camera_pos(frame_index),
                                first_ray=first, n_rays=count)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


_FRAME0 = None


def cpu_baseline_sample(target=20480):
    """Bounded sample of the same workload for the CPU arm: a strided subsample of frame 0 with about
    `target` rays (stride 257 -> 20 401 rays), so that it covers the whole image including the shadow."""
    global _FRAME0
    if _FRAME0 is None:
        _FRAME0 = frame_rays(0)
    pos, d = _FRAME0
    stride = max(1, pos.shape[0] // max(int(target), 1)) | 1   # odd: coprime with the image width, so the sample
    # walks diagonally through the image instead of picking 4 columns (one of them the pole-grazing centre line)
    sel = np.arange(0, pos.shape[0], stride)
    return pos[sel], d[sel], stride


def raygen_r_sphere():
    from blackhole_geodesic_calculator_b200 import raygen
    return float(raygen.CFG_R_SPHERE)


def time_reference(pos, d, processes):
    """The reference's method (sympy RHS + scipy.solve_ivp RK45, one Python call per ray) on host cores."""
    import multiprocessing as mp
    from oracle import schwarzschild_ref as R
    R._build_rhs()  # the sympy derivation happens once per frame in the reference (RRE.py:134): not timed
    pool = mp.get_context("fork").Pool(processes)
    try:
        R.trace_pool(pos[:processes * 8], d[:processes * 8], processes, chunk=8, pool=pool)  # warm the workers
        t0 = time.perf_counter()
        out = R.trace_pool(pos, d, processes, chunk=max(16, min(256, pos.shape[0] // (4 * processes))), pool=pool)
        dt = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    return pos.shape[0] / dt, dt, out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # size the per-step sample so that steps+warmup finish in a few minutes: ~8 s of all-core work per step
    pos, d, _ = cpu_baseline_sample(max(cores * 16, 256))
    probe_rate, _, _ = time_reference(pos, d, cores)
    pos, d, stride = cpu_baseline_sample(min(20480, max(cores * 16, probe_rate * 8.0)))
    rates = []
    per_step = pos.shape[0]
    t_all = 0.0
    for i in range(args.warmup + args.steps):
        r, dt, _ = time_reference(pos, d, cores)
        if i >= args.warmup:
            rates.append(r)
            t_all += dt
    value = per_step * len(rates) / t_all
    line = {
        "impl": "reference", "metric": "geodesic rays/s (1024^2 x 5 spp Schwarzschild frame)", "value": value,
        "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_all / len(rates), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config 2: 1024x1024x5spp Schwarzschild frame, M=1, r_sphere=60M, rtol=1e-3, "
                               "atol=1e-6 (bounded sample)", "rays_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"every {stride}th ray of frame 0 ({per_step} rays per step); restated "
                                   "reference method (sympy RHS + scipy.solve_ivp RK45 per ray, multiprocessing "
                                   "Pool over all cores); curvedpy itself is absent"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def strong_frame(args, dev, local, rank, world, n1_ms):
    """ONE config-2 frame sharded over all ranks and delivered into rank 0's HBM (the gather to the rank
    that owns the Blender frame).  Untimed: bit-equality of both multi-GPU routes with the single-GPU call on the same
    rays.  Timed: CUDA events around the sharded call, barrier + synchronize on both sides, max over ranks."""
    import torch
    import torch.distributed as dist
    from blackhole_geodesic_calculator_b200 import api, distributed as D, raygen

    n = W * H * SPP
    cpos = frame_camera_pos(0)
    cam = api.make_camera(cpos, raygen.look_at_rotation(cpos), W, H * SPP, raygen.CFG_FOV, raygen.CFG_FOV,
                          seed=raygen.CFG_SEED, jitter="philox")
    cam.height = H
    pos0, d0, _ = api.generate_rays(cam, n, raygen.CFG_R_SPHERE, device=local)  # identical on every rank
    kw = dict(M=raygen.CFG_M, r_sphere=raygen.CFG_R_SPHERE, rtol=1e-3, atol=1e-6, mode=args.mode)
    bits = lambda t: t.contiguous().view(torch.int64 if t.dtype == torch.float64 else torch.int32)
    same = lambda a, b: all(torch.equal(bits(x), bits(y)) for x, y in zip(a, b))
    single = api.trace(pos0, d0, image_width=W, **kw)[:3]
    gathered = D.trace_sharded(pos0, d0, **kw)
    out = {"rays": n, "n1_ms": n1_ms, "nvlink_bytes_per_frame": (world - 1) * (n // world) * 52}
    if rank == 0:
        out["gather_equals_single"] = same(gathered, single)
    frame = D.PeerFrame(n, owner=0)
    try:
        def timed(fn, iters):
            ts = []
            for it in range(iters + 2):
                dist.barrier()
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                dist.barrier()
                torch.cuda.synchronize(dev)
                t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if it >= 2:
                    ts.append(float(t))
            return float(np.median(ts)), float(np.min(ts))

        iters = max(3, min(args.steps, 10))
        routes = {}
        equal = True
        variants = [("courier", "flags", W, 1), ("stores", "flags", 0, 1), ("copy", "flags", W, 2), ("courier", "nccl", W, 1)]
        for route, sync, width, chunks in variants:
            if rank == 0:
                frame.tensors()[0].fill_(float("nan"))
                frame.tensors()[1].fill_(float("nan"))
                frame.tensors()[2].fill_(-7)
                torch.cuda.synchronize(dev)
            dist.barrier()
            call = lambda: D.trace_sharded_peer(pos0, d0, frame, image_width=width, route=route, sync=sync,
                                                chunks=chunks, **kw)
            res = call()
            torch.cuda.synchronize(dev)
            dist.barrier()
            if rank == 0:
                equal = equal and same(res, single)
            med, best = timed(call, iters)
            routes[f"{route}/{sync}/w{width}/c{chunks}"] = {"ms_median": med, "ms_min": best}
        # floor: the same shard integrated into LOCAL buffers (no delivery at all)
        order = frame.order(0)
        lp, ld = torch.empty_like(pos0), torch.empty_like(d0)
        ls = torch.empty(n, dtype=torch.int32, device=dev)
        prm = api.make_params(**kw)
        local_call = lambda: api.trace_device(pos0.data_ptr(), d0.data_ptr(), lp.data_ptr(), ld.data_ptr(), ls.data_ptr(),
                                              None, order.data_ptr(), order.numel(), api.LAYOUT_AOS, prm, local,
                                              torch.cuda.current_stream(dev).cuda_stream)
        local_call()
        out["shard_compute_only_ms"] = timed(local_call, iters)[0]
        g_med, _ = timed(lambda: D.trace_sharded(pos0, d0, **kw), iters)
        best_route = min(routes, key=lambda r: routes[r]["ms_median"])
        t_ms = routes[best_route]["ms_median"]
        out.update({"route": best_route, "ms": t_ms, "value": n / (t_ms * 1e-3), "unit": "rays/s",
                    "efficiency_vs_n1": (n1_ms / (world * t_ms)) if n1_ms else None,
                    "ms_by_route": routes, "nccl_gather_route_ms": g_med, "peer_equals_single": equal if rank == 0 else None,
                    "what": "one 1024x1024x5spp frame: 8-row bands dealt cyclically to the ranks; routes: courier = one "
                            "trace launch per rank whose finished bands a 2-SM courier kernel stores into rank 0's HBM "
                            "over NVLink while the integration runs, stores = the trace kernel's own (staged) remote "
                            "stores, copy = pieces + copy-engine deliveries; arrival flags by stream memory operations "
                            "(or an NCCL fence); timed from the call to the owner's stream having every shard (max "
                            "over ranks)"})
    finally:
        frame.close()
    return out


def other_configs(dev, local, peak_tf):
    """BASELINE configs 3 and 5 on the same build (parity-test cases; reported here so that the driver's run holds them):
    device-resident rays, CUDA events, median of 5.  Config 5 in random planes has no image order: it goes through the
    cost binning."""
    import torch
    from blackhole_geodesic_calculator_b200 import api, raygen

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    out = {}
    p3, d3 = raygen.random_impact_bundle(None)
    bundles = [("config3_1920x1080_random_impact", p3, d3)]
    for name, inplane in (("config5_near_critical_random_planes", False), ("config5_near_critical_in_plane", True)):
        p5, d5, _ = raygen.near_critical_bundle(1 << 20, in_plane=inplane)
        bundles.append((name, p5, d5))
    for name, p, d in bundles:
        tp, td = torch.from_numpy(p).to(dev), torch.from_numpy(d).to(dev)
        _, _, st, cnt = api.trace(tp, td, return_counters=True)
        att = float(cnt[0].double().sum())
        integ = float((st != 2).double().sum())
        ms = timed(lambda: api.trace(tp, td))
        flop = FLOP_FIXED * integ + FLOP_PER_ATTEMPT * att
        tf = flop / (ms * 1e-3) / 1e12
        out[name] = {"rays": int(p.shape[0]), "ms": ms, "rays_per_s": p.shape[0] / (ms * 1e-3),
                     "attempts_per_ray": att / p.shape[0], "tflops": tf, "frac_of_peak": tf / peak_tf if peak_tf else None}
        del tp, td
    return out


def animation_100(args, dev, local, rank, world):
    """Config 4: 100 distinct frames of the orbiting camera (azimuth +3.6 deg per frame), 5 242 880 rays each.  The
    first (100 // N) N frames go to rank f mod N (distributed.frames_for_rank) and are traced straight from their
    176-byte camera description into alternating device buffers (no host ray buffers, no collective); the remaining
    100 mod N frames are each split by 8-row bands over ALL ranks and delivered into rank 0's HBM (courier route)."""
    import ctypes
    import torch
    import torch.distributed as dist
    from blackhole_geodesic_calculator_b200 import _lib, api, distributed as D, raygen

    n, frames = W * H * SPP, 100
    lib = _lib.load()
    params = api.make_params(M=raygen.CFG_M, r_sphere=raygen.CFG_R_SPHERE, rtol=1e-3, atol=1e-6, mode=args.mode)
    stream = torch.cuda.current_stream(dev).cuda_stream
    bufs = [(torch.empty((n, 3), dtype=torch.float64, device=dev), torch.empty((n, 3), dtype=torch.float64, device=dev),
             torch.empty(n, dtype=torch.int32, device=dev)) for _ in range(2)]

    def camera(f):
        cpos = frame_camera_pos(f)
        cam = api.make_camera(cpos, raygen.look_at_rotation(cpos), W, H * SPP, raygen.CFG_FOV, raygen.CFG_FOV,
                              seed=raygen.CFG_SEED, jitter="philox")
        cam.height = H
        return cam

    full = (frames // world) * world
    mine = [camera(f) for f in D.frames_for_rank(full, rank, world)]
    left = [camera(f) for f in range(full, frames)]
    frame, split = None, bool(left and world > 1)
    if split:
        try:
            frame = D.PeerFrame(n, owner=0)
        except Exception:   # no peer mapping on this box: the leftover frames go to single ranks instead of being split
            split = False
        ok = torch.tensor([1 if split else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok) and frame is not None:
            frame.close()
            frame = None
        split = bool(int(ok))
    if left and not split:
        mine += [cam for i, cam in enumerate(left) if i % world == rank]
        left = []
    kw = dict(M=raygen.CFG_M, r_sphere=raygen.CFG_R_SPHERE, rtol=1e-3, atol=1e-6, mode=args.mode)

    def run():
        for i, cam in enumerate(mine):
            ep, ed, st = bufs[i & 1]
            _lib.check(lib.bhg_trace_camera_f64(ctypes.byref(cam), ep.data_ptr(), ed.data_ptr(), st.data_ptr(), None, n,
                                                ctypes.byref(params), local, stream or None))
        for cam in left:
            pos, d, _ = api.generate_rays(cam, n, raygen.CFG_R_SPHERE, device=local)
            D.trace_sharded_peer(pos, d, frame, image_width=W, route="courier", **kw)

    try:
        run()   # warm-up (allocator pools, peer mappings)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
    finally:
        if frame is not None:
            frame.close()
    ms = float(t)
    return {"frames": frames, "rays": frames * n, "ms": ms, "value": frames * n / (ms * 1e-3), "unit": "rays/s",
            "frames_per_rank": len(mine), "leftover_frames_split_by_bands": len(left) if split or world == 1 else 0,
            "input_bytes_per_frame": 176, "results": "device-resident (two alternating buffer sets per rank)",
            "ms_per_frame_equivalent": ms / frames}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from blackhole_geodesic_calculator_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = W * H * SPP
    pos_h, dir_h = frame_rays(rank)  # frame index = rank: frames shard across GPUs
    # pinned host buffers for the end-to-end leg
    pin_pos, pin_dir = api.pinned_empty((n, 3), device=local), api.pinned_empty((n, 3), device=local)
    pin_pos[:] = pos_h
    pin_dir[:] = dir_h
    pin_op, pin_od = api.pinned_empty((n, 3), device=local), api.pinned_empty((n, 3), device=local)
    pin_st = api.pinned_empty((n,), np.int32, device=local)
    pos = torch.from_numpy(pos_h).to(dev)
    d = torch.from_numpy(dir_h).to(dev)
    exit_pos, exit_dir = torch.empty_like(pos), torch.empty_like(pos)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    counters = torch.empty((2, n), dtype=torch.int32, device=dev)
    params = api.make_params(mode=args.mode, refill_threshold=args.threshold, image_width=0 if args.no_tiles else W)
    stream = torch.cuda.current_stream(dev)

    extras = None
    if args.disk:  # opt-in in-flight disk-crossing event (next-row 2), annulus 6 M .. 20 M
        from blackhole_geodesic_calculator_b200._lib import BhgExtras
        disk_xy = torch.empty((n, 2), dtype=torch.float64, device=dev)
        extras = BhgExtras(6.0, 20.0, disk_xy.data_ptr())

    def step(with_counters=False):
        api.trace_device(pos.data_ptr(), d.data_ptr(), exit_pos.data_ptr(), exit_dir.data_ptr(), status.data_ptr(),
                         counters.data_ptr() if with_counters else None, None, n, api.LAYOUT_AOS, params, local,
                         stream.cuda_stream, extras)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # attempts per ray (kernel's own counters) for the algorithmic-FLOP figure; untimed
    step(with_counters=True)
    torch.cuda.synchronize(dev)
    att, acc, integ = api.sum_counters(counters.data_ptr(), status.data_ptr(), n, local, stream.cuda_stream)
    if args.mode == "plane":
        flop_per_launch = FLOP_FIXED_PLANE * integ + FLOP_PER_ATTEMPT_PLANE * att
    else:
        flop_per_launch = FLOP_FIXED * integ + FLOP_PER_ATTEMPT * att

    # clocks are sampled from before the warm-up to the end of the timed region (same load throughout); the warm-up
    # runs at least W steps and at least 0.4 s so that nvidia-smi is up and several samples fall under load
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_w = time.perf_counter()
    n_w = 0
    while n_w < max(args.warmup, 3) or time.perf_counter() - t_w < 0.4:
        step()
        n_w += 1
        if n_w % 8 == 0:
            torch.cuda.synchronize(dev)
    barrier()
    launches0 = api.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for e0, e1 in evs:
        e0.record(stream)
        step()
        e1.record(stream)
    t_end.record(stream)
    barrier()
    launches = api.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    kern_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end through the public host API: pinned numpy in -> H2D -> trace -> D2H -> pinned numpy out
    lib_params = params
    from blackhole_geodesic_calculator_b200 import _lib
    import ctypes
    lib = _lib.load()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)

    def e2e_step():
        _lib.check(lib.bhg_trace_schwarzschild_f64_host(p(pin_pos), p(pin_dir), p(pin_op), p(pin_od), p(pin_st), None,
                                                        n, ctypes.byref(lib_params), local))

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    # end-to-end from the camera description: ~150 bytes in, rays generated on the device,
    # (a) all outputs, (b) directions + status only (what the RRE / CAM consumers read)
    from blackhole_geodesic_calculator_b200 import raygen
    cpos = frame_camera_pos(rank)
    cam = api.make_camera(cpos, raygen.look_at_rotation(cpos), W, H * SPP, raygen.CFG_FOV, raygen.CFG_FOV,
                          seed=raygen.CFG_SEED, jitter="philox")
    # the s -> y -> x stack of SPP images is addressed as one image of H*SPP rows: same rays, same order
    cam.height = H
    cam_times = []
    for want_pos in (True, False):
        bufs = (pin_op if want_pos else None, pin_od, pin_st)
        api.trace_camera(cam, n, mode=args.mode, refill_threshold=args.threshold, device=local, want_pos=want_pos,
                         buffers=bufs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            api.trace_camera(cam, n, mode=args.mode, refill_threshold=args.threshold, device=local,
                             want_pos=want_pos, buffers=bufs)
        torch.cuda.synchronize(dev)
        cam_times.append(time.perf_counter() - t0)

    # (c) camera -> sky-lookup coordinates (next-row 3): 12 B/ray come back
    pin_uv = api.pinned_empty((n, 2), np.float32, device=local)
    api.trace_camera_sky(cam, n, mode=args.mode, refill_threshold=args.threshold, device=local, buffers=(pin_uv, pin_st))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        api.trace_camera_sky(cam, n, mode=args.mode, refill_threshold=args.threshold, device=local,
                             buffers=(pin_uv, pin_st))
    torch.cuda.synchronize(dev)
    cam_times.append(time.perf_counter() - t0)

    # (d) the same host-buffer contract with float32 arrays (Blender's native precision): 28 B/ray over PCIe
    f_pos, f_dir = api.pinned_empty((n, 3), np.float32, device=local), api.pinned_empty((n, 3), np.float32, device=local)
    f_pos[:] = pos_h
    f_dir[:] = dir_h
    f_out = (api.pinned_empty((n, 3), np.float32, device=local), api.pinned_empty((n, 3), np.float32, device=local), pin_st)
    kw32 = dict(mode=args.mode, refill_threshold=args.threshold, image_width=0 if args.no_tiles else W, device=local,
                out=f_out)
    api.trace_f32(f_pos, f_dir, **kw32)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        api.trace_f32(f_pos, f_dir, **kw32)
    torch.cuda.synchronize(dev)
    cam_times.append(time.perf_counter() - t0)

    # (e) camera in, float32 directions + status out: 16 B/ray over PCIe (the recommended wiring for the RRE / CAM
    # consumers, which read exit_dir only)
    f_dir16 = f_out[1]
    if args.mode == "parity":
        api.trace_camera_f32(cam, n, device=local, buffers=(None, f_dir16, pin_st))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            api.trace_camera_f32(cam, n, device=local, buffers=(None, f_dir16, pin_st))
        torch.cuda.synchronize(dev)
        cam_times.append(time.perf_counter() - t0)
    else:
        cam_times.append(float("nan"))

    t = torch.tensor([total_ms, e2e_s * 1e3] + [c * 1e3 for c in cam_times], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, cam_full_ms, cam_dir_ms, cam_uv_ms, f32_ms, cam_f32_ms = (float(v) for v in t)
    value = world * n * args.steps / (total_ms * 1e-3)
    e2e_value = world * n * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        kavg = float(np.mean(kern_ms))
        try:
            peak_tf, clk_est = api.fp64_peak_tflops(local)
        except Exception:
            peak_tf, clk_est = None, None
        achieved_tf = flop_per_launch / (kavg * 1e-3) / 1e12
        hbm_peak = 6650.0
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak = float(json.load(f).get("hbm_gbs", hbm_peak))
        except Exception:
            pass
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get(args.mode, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        peak = peak_tf or NOMINAL_FP64_TFLOPS
        line = {
            "metric": "geodesic rays/s (1024^2 x 5 spp Schwarzschild frame)", "value": value, "unit": "rays/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "warmup_steps_run": n_w,
            "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config 2: 1024x1024x5spp Schwarzschild frame (5242880 rays/GPU/step), M=1, "
                                   "r_sphere=60M, rtol=1e-3, atol=1e-6, camera (120,-80,40)M fov 0.6, Philox jitter "
                                   "seed 42; at N>1 rank r integrates animation frame r (config 4, camera azimuth "
                                   "+3.6 deg/frame)",
                       "mode": args.mode, "refill_threshold": args.threshold or "adaptive (idle budget 96 lane-iterations)",
                       "image_width_hint": 0 if args.no_tiles else W, "disk_event": bool(args.disk),
                       "rays_per_step_per_gpu": n, "cache": "inputs+outputs 525 MB per step > 126 MB L2 (no flush)",
                       "mean_attempts_per_ray": att / n,
                       "layout": "float64 AoS entry/exit arrays; the trace kernel reads prepared records (pre-pass) as "
                                 "coalesced double2 planes in queue order; SoA arrays measured 0.8 % slower with the tile "
                                 "order and equal without it, bit-identical results (profiles/r2k_layout_probe.json)",
                       "pre_pass": "prepare_kernel (entry conversion, f0, Hairer step) is inside the timed kernel_ms"},
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": n * 48, "d2h_bytes_per_step": n * 52,
                    "steps": e2e_steps, "api": "bhg_trace_schwarzschild_f64_host (pinned numpy in/out, chunked "
                                               "H2D/trace/D2H pipeline)"},
            "e2e_camera": {"all_outputs": {"value": world * n * e2e_steps / (cam_full_ms * 1e-3), "unit": "rays/s",
                                           "h2d_bytes_per_step": 176, "d2h_bytes_per_step": n * 52},
                           "dir_and_status": {"value": world * n * e2e_steps / (cam_dir_ms * 1e-3), "unit": "rays/s",
                                              "h2d_bytes_per_step": 176, "d2h_bytes_per_step": n * 28},
                           "sky_uv_and_status": {"value": world * n * e2e_steps / (cam_uv_ms * 1e-3), "unit": "rays/s",
                                                 "h2d_bytes_per_step": 176, "d2h_bytes_per_step": n * 12},
                           "dir_f32_and_status": {"value": world * n * e2e_steps / (cam_f32_ms * 1e-3), "unit": "rays/s",
                                                  "h2d_bytes_per_step": 176, "d2h_bytes_per_step": n * 16,
                                                  "api": "bhg_trace_camera_f32_host"},
                           "api": "bhg_trace_camera_f64_host / bhg_trace_camera_sky_host: rays generated on the device from the camera struct "
                                  "(next-row 1), pinned numpy outputs"},
            "e2e_f32io": {"value": world * n * e2e_steps / (f32_ms * 1e-3), "unit": "rays/s",
                          "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 28,
                          "api": "bhg_trace_schwarzschild_f32io_host: float32 [N,3] host arrays in/out (Blender's native "
                                 "precision), FP64 integration"},
            "gpu_launches": int(launches),
            "kernel_ms": {"mean": kavg, "min": float(np.min(kern_ms)), "max": float(np.max(kern_ms))},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak,
                         "peak_source": ("on-box DFMA microbenchmark (MEASURED_PEAKS.json has no FP64 entry); "
                                         f"nominal {NOMINAL_FP64_TFLOPS}") if peak_tf else "nominal (148 SM x 64 FMA/clk x 2 x 1.965 GHz)",
                         "frac_of_nominal": achieved_tf / NOMINAL_FP64_TFLOPS,
                         "algorithmic_flop_per_launch": flop_per_launch,
                         "hbm_gbs_achieved": n * 100 / (kavg * 1e-3) / 1e9,
                         "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write, profiles/ncu_traffic.json)",
                         "traffic_note": "above the algorithmic 100 B/ray on purpose: the pre-pass (0.2 ms, 717 MB) turns "
                                         "the 48 B/ray entry state into a 96 B/ray prepared record that the trace kernel "
                                         "reads as coalesced double2 planes; HBM stays at 3 % of its bandwidth, the FP64 "
                                         "pipe is the bound",
                         "algorithmic_bytes_per_launch": n * 100, "dfma_clock_mhz_est": clk_est},
            "roofline_hbm": {"bound": "hbm", "achieved": n * 100 / (kavg * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": n * 100 / (kavg * 1e-3) / 1e9 / hbm_peak,
                             "note": "secondary: 100 B/ray of algorithmic traffic against the measured copy bandwidth "
                                     "(MEASURED_PEAKS.json hbm_gbs, else fallback 6650) - shows HBM is negligible"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cpos, cdir, stride = cpu_baseline_sample(min(20480, max(256, 250 * cores * 12)))  # ~12 s of work
            m = cpos.shape[0]
            rate, dt, cpu_out = time_reference(cpos, cdir, cores)
            line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                                    "sample": f"every {stride}th ray of frame 0 ({m} rays) in {dt:.1f} s; "
                                              "restated reference method (sympy RHS + scipy.solve_ivp RK45 per ray, "
                                              "multiprocessing Pool over all cores)"}
            if args.mode == "parity" and not args.disk:
                try:
                    # the CPU arm just integrated every `stride`-th ray of the frame the GPU traced: compare them
                    sel = torch.arange(0, n, stride, device=dev)
                    g_pos, g_dir = exit_pos.index_select(0, sel).cpu().numpy(), exit_dir.index_select(0, sel).cpu().numpy()
                    g_st = status.index_select(0, sel).cpu().numpy()
                    g_acc = counters[1].index_select(0, sel).cpu().numpy()
                    c_pos, c_dir, c_st, _, c_acc = cpu_out[:5]
                    esc = (c_st == 0) & (g_st == 0)
                    nrm = np.cross(cpos, cdir)
                    pole = np.abs(nrm[:, 2]) / np.linalg.norm(nrm, axis=1) < 1e-2   # orbital plane contains the polar axis
                    dpos = np.abs(g_pos - c_pos).max(axis=1) / raygen_r_sphere()
                    ddir = np.abs(g_dir - c_dir).max(axis=1)
                    line["parity_vs_cpu_sample"] = {
                        "rays": int(m), "status_equal": bool((g_st == c_st).all()),
                        "accepted_steps_equal_frac": float((g_acc == c_acc).mean()),
                        "max_rel_dpos_escaped": float(dpos[esc & ~pole].max()),
                        "max_ddir_escaped": float(ddir[esc & ~pole].max()),
                        "pole_grazing": {"rays": int((esc & pole).sum()), "over_1e-6": int((dpos[esc & pole] > 1e-6).sum()),
                                         "max_rel_dpos": float(dpos[esc & pole].max(initial=0.0))},
                        "note": "GPU results of the timed frame against the scipy arm on the same rays; rays whose "
                                "orbital plane contains the polar axis (|n_z| < 1e-2) cross the coordinate singularity "
                                "of the reference's formulation and are listed apart (DESIGN.md section 2); "
                                "informational - the parity gate is tests/ (pytest -m gpu)"}
                except Exception as e:  # informational block: never lose the bench line over it
                    line["parity_vs_cpu_sample"] = {"error": str(e)}
            try:
                # the two other figures SURVEY.md section 8(d) asks for: one process, and the RRE call shape
                # (no sphere, lambda in [0, 50], 10 000 t_eval samples per ray of which the engine reads the last:
                # RelativisticRenderEngine.py:293-294,307-308)
                from oracle import schwarzschild_ref as R
                from blackhole_geodesic_calculator_b200 import raygen
                k1 = min(512, m)
                t0 = time.perf_counter()
                R.trace(cpos[:k1], cdir[:k1], 1.0, raygen.CFG_R_SPHERE, 1e-3, 1e-6)
                line["cpu_baseline"]["single_process_rays_per_s"] = k1 / (time.perf_counter() - t0)
                rot = raygen.look_at_rotation((12.0, -8.0, 4.0))
                rd = raygen.camera_rays(16, 16, 1, 1.0, 1.0, rot, 42, "philox")
                rp = np.tile([12.0, -8.0, 4.0], (rd.shape[0], 1))
                t0 = time.perf_counter()
                R.trace_pool(rp, rd, cores, chunk=max(1, rd.shape[0] // (2 * cores)), M=0.5, r_sphere=np.inf,
                             lambda_max=50.0, polyline=10000)
                line["cpu_baseline"]["rre_call_shape_rays_per_s"] = rd.shape[0] / (time.perf_counter() - t0)
                line["cpu_baseline"]["rre_call_shape_sample"] = ("256 rays from a camera at (12,-8,4) M, M=0.5, no "
                                                                 "sphere, curve_end=50, nr_points_curve=10000, Pool "
                                                                 "over all cores (includes starting the pool)")
            except Exception as e:
                line["cpu_baseline"]["extra_shapes_error"] = str(e)
            if args.mode == "parity":
                try:
                    line["other_configs"] = other_configs(dev, local, peak)
                except Exception as e:
                    line["other_configs"] = {"error": repr(e)[:300]}
                try:
                    # the same live check against the REAL scipy arm on samples of configs 3 and 5 (VERDICT r1 task 1)
                    from blackhole_geodesic_calculator_b200 import raygen
                    extra = {}
                    p3, d3 = raygen.random_impact_bundle(None)
                    sel3 = np.arange(0, p3.shape[0], 199)
                    p5, d5, _ = raygen.near_critical_bundle(1 << 20, in_plane=False)
                    sel5 = np.arange(0, 1 << 20, 129)
                    for name, pp, dd in (("config3_sample", p3[sel3], d3[sel3]), ("config5_random_planes_sample", p5[sel5], d5[sel5])):
                        _, _, c_out = time_reference(pp, dd, cores)
                        g = api.trace(pp, dd, return_counters=True)
                        c_pos, c_dir, c_st, _, c_acc = c_out[:5]
                        same = (g[3][1] == c_acc)
                        esc = (c_st == 0) & (g[2] == 0)
                        dev_ = np.maximum(np.abs(g[0] - c_pos).max(axis=1) / raygen_r_sphere(), np.abs(g[1] - c_dir).max(axis=1))
                        extra[name] = {"rays": int(pp.shape[0]), "status_equal": bool((g[2] == c_st).all()),
                                       "accepted_steps_equal_frac": float(same.mean()),
                                       "escaped_beyond_1e-6": int((dev_[esc] > 1e-6).sum()),
                                       "escaped_beyond_1e-6_with_equal_steps": int((dev_[esc & same] > 1e-6).sum()),
                                       "median_dev_escaped": float(np.median(dev_[esc])) if esc.any() else None,
                                       "max_dev_escaped": float(dev_[esc].max(initial=0.0))}
                    extra["note"] = ("GPU against the scipy arm on every 199th ray of config 3 and every 129th of config 5 in random "
                                     "planes; near-critical pole-grazing rays of config 5 are ill-conditioned in the reference's "
                                     "own formulation (scipy vs its C restatement disagree on the same rays, "
                                     "profiles/r2a_adjudication.json): the gate is the per-ray conditioning rule of tests/")
                    line["parity_vs_cpu_sample_other_configs"] = extra
                except Exception as e:
                    line["parity_vs_cpu_sample_other_configs"] = {"error": repr(e)[:300]}
            try:
                from oracle import port
                t0 = time.perf_counter()
                port.trace(cpos, cdir)
                line["cpu_baseline"]["c_port_rays_per_s_all_cores"] = cpos.shape[0] / (time.perf_counter() - t0)
            except Exception as e:  # the C port is optional information
                line["cpu_baseline"]["c_port_error"] = str(e)
    strong = None
    if world > 1 and args.mode == "parity" and not args.no_strong:
        n1 = torch.tensor([float(np.mean(kern_ms)) if rank == 0 else 0.0], dtype=torch.float64, device=dev)
        dist.broadcast(n1, src=0)
        try:
            strong = strong_frame(args, dev, local, rank, world, float(n1))
        except Exception as e:  # never lose the bench line over the extra section
            strong = {"error": repr(e)[:400]}
    anim = None
    if args.mode == "parity" and not args.no_animation:
        try:
            anim = animation_100(args, dev, local, rank, world)
        except Exception as e:
            anim = {"error": repr(e)[:400]}
    if rank == 0:
        if anim is not None:
            line["animation_100"] = anim
        if strong is not None:
            line["strong_frame"] = strong
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="parity", choices=["parity", "plane"])
    ap.add_argument("--threshold", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tiles", action="store_true", help="do not pass the image_width scheduling hint")
    ap.add_argument("--no-animation", action="store_true", help="skip the 100-frame animation section (config 4)")
    ap.add_argument("--no-strong", action="store_true", help="skip the single-frame strong-scaling section at N > 1")
    ap.add_argument("--disk", action="store_true", help="also locate equatorial-disk crossings (6 M .. 20 M) in flight")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
