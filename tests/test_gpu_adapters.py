"""GPU tests of the reference-shaped adapters (RRE / LIM / CAM call shapes) against the oracle."""
import numpy as np
import pytest

from conftest import assert_parity, golden_kwargs, load_golden

pytestmark = pytest.mark.gpu


def test_rre_call_shape_per_ray_and_batch():
    from blackhole_geodesic_calculator_b200 import adapters
    g = load_golden("rre_shape_16x16.npz")
    kw = golden_kwargs(g)
    gi = adapters.GeodesicIntegratorSchwarzschild(mass=kw["M"], time_like=False)
    # batched body of spacetime_ray_cast (RRE.py:271-313): camera at `origin`, BH at bh_loc
    bh = np.array([1.0, 2.0, 3.0])
    origin = g["entry_pos"][0] + bh
    hit, hit_bh, end_dir, end_loc = adapters.spacetime_ray_cast_batch(gi, origin, g["entry_dir"], bh_loc=bh,
                                                                    curve_end=kw["lambda_max"])
    assert not hit.any()
    assert np.array_equal(hit_bh, g["status"] == 1)
    ok = g["status"] == 3  # the normal RRE ending: affine length used up
    assert np.abs(end_loc[ok] - g["exit_pos"][ok]).max() / 50.0 < 1e-6
    assert np.abs(end_dir[ok] - g["exit_dir"][ok]).max() < 1e-6
    # per-ray drop-in with the reference's return shape (RRE.py:293-297,307-308)
    for i in (0, 17, 200):
        k_xyz, x_xyz, result = gi.calc_trajectory(g["entry_dir"][i], g["entry_pos"][i], curve_end=kw["lambda_max"],
                                                  nr_points_curve=10000)
        assert result["start_inside_hole"] is False
        assert result["hit_blackhole"] == bool(g["status"][i] == 1)
        x, y, z = x_xyz
        kx, ky, kz = k_xyz
        assert x_xyz.shape[1] > 2 and np.allclose(x_xyz[:, 0], g["entry_pos"][i])
        end_loc_i = np.array([x[-1], y[-1], z[-1]])
        end_dir_i = np.array([kx[-1], ky[-1], kz[-1]])
        assert np.abs(end_loc_i - g["exit_pos"][i]).max() / 50.0 < 1e-6
        assert np.abs(end_dir_i - g["exit_dir"][i]).max() < 1e-6
    with pytest.raises(NotImplementedError):
        adapters.GeodesicIntegratorSchwarzschild(time_like=True)


def test_lim_call_shape_scaling_and_messages():
    from blackhole_geodesic_calculator_b200 import adapters, raygen
    from oracle import port
    sw = adapters.SchwarzschildGeodesic(metric="schwarzschild", coordinates="schwarzschild")
    ratio, r_obj = 30.0, 3.0            # a Blender sphere of radius 3 standing for 30 r_s (LIM.py:488)
    pos, d = raygen.config_bundle(24, 24, 1, fov=0.5, r_sphere=60.0)
    locs = pos / 60.0 * r_obj           # entry points on the Blender sphere, relative to its centre
    end_loc, end_dir, hit_bh, outside, status = sw.ray_trace_batch(d, locs, exit_tolerance=0.2,
                                                                   ratio_obj_to_blackhole=ratio)
    o = port.trace(locs * (ratio / r_obj), d, M=0.5, r_sphere=ratio, lambda_max=sw.approximateCurveEnd(ratio))
    assert np.array_equal(status, o["status"])
    assert_parity(end_loc * (ratio / r_obj), end_dir, status, o["exit_pos"], o["exit_dir"], o["status"], ratio)
    assert np.array_equal(hit_bh, o["status"] == 1) and not outside.any()
    # exit points are back on the Blender sphere
    assert np.abs(np.linalg.norm(end_loc[~hit_bh], axis=1) - r_obj).max() < 1e-9
    # per-ray shape and the 'Outside' message when the affine length runs out (LIM.py:308-314)
    x, y, z, el, ed, mes = sw.ray_trace(d[5], locs[5], ratio_obj_to_blackhole=ratio)
    assert mes["hit_blackhole"] == bool(o["status"][5] == 1) and "error" not in mes
    assert np.allclose(el, end_loc[5]) and len(x) > 10
    # the polyline (r_s units) starts at the entry point, stays inside the sphere and ends at the exit point
    rr = np.sqrt(x * x + y * y + z * z)
    assert abs(rr[0] - ratio) < 1e-9 and abs(rr[-1] - ratio) < 1e-6 and (rr <= ratio + 1e-9).all()
    # the reference's own checkHitDisk scan (LIM.py:413-438) applied to that polyline agrees with the in-flight event
    def check_hit_disk(x, y, z, R_in, R_out):
        for i in range(len(x) - 1):
            if (z[i + 1] < 0 and z[i] >= 0) or (z[i + 1] > 0 and z[i] <= 0):
                l0 = -z[i] / (z[i + 1] - z[i])
                xd, yd = x[i] + (x[i + 1] - x[i]) * l0, y[i] + (y[i + 1] - y[i]) * l0
                R = np.sqrt(xd ** 2 + yd ** 2)
                if R_in <= R <= R_out:
                    return np.array([xd, yd])
        return None
    res = sw.ray_trace_batch(d, locs, ratio_obj_to_blackhole=ratio, disk=(3.0, 12.0))
    n_hit = 0
    for i in np.nonzero(np.isfinite(res[-1][:, 0]))[0][:6]:
        xi, yi, zi, *_ = sw.ray_trace(d[i], locs[i], ratio_obj_to_blackhole=ratio, nr_points_curve=2048)
        h = check_hit_disk(xi, yi, zi, 3.0, 12.0)
        assert h is not None and np.abs(h - res[-1][i]).max() < 0.05   # chord interpolation vs exact crossing
        n_hit += 1
    assert n_hit > 0
    x, y, z, el, ed, mes = sw.ray_trace(d[5], locs[5], ratio_obj_to_blackhole=ratio, curve_end=5.0)
    assert mes.get("error") == "Outside" and mes["hit_blackhole"] is False
    # the engine's `approx` branch (LIM.py:97-101,269): same attributes, same call, answered exactly
    asw = adapters.ApproxSchwarzschildGeodesic(ratio_obj_to_blackhole=ratio, exit_tolerance=0.2,
                                               coordinates="schwarzschild")
    assert round(asw.exit_tolerance, 4) == 0.2 and round(asw.ratio_obj_to_blackhole, 4) == ratio
    a_loc, a_dir, a_mes = asw.generatedRayTracer(locs[5], d[5])
    assert np.array_equal(a_loc, end_loc[5]) and np.array_equal(a_dir, end_dir[5])
    assert a_mes["hit_blackhole"] == bool(hit_bh[5]) and "error" not in a_mes
    cap = int(np.nonzero(hit_bh)[0][0])
    assert asw.generatedRayTracer(locs[cap], d[cap])[2]["hit_blackhole"] is True
    # default chart of this call shape: isotropic (README Fig. 5 / 6, tests/test_readme_figures.py) - the same rays
    # read in that chart, against the oracle wrapped in the host-side maps (exit state, disk hit point, polyline)
    from blackhole_geodesic_calculator_b200 import coords as C
    swi = adapters.SchwarzschildGeodesic()
    assert swi.coordinates == "isotropic"
    end_loc, end_dir, hit_bh, outside, status, dxy = swi.ray_trace_batch(d, locs, ratio_obj_to_blackhole=ratio,
                                                                         disk=(3.0, 12.0))
    ps, ds = C.isotropic_to_schwarzschild(locs * (ratio / r_obj), d, 0.5)
    R_s = float(C.schwarzschild_radius(ratio, 0.5))
    o = port.trace(ps, ds, M=0.5, r_sphere=R_s, lambda_max=swi.approximateCurveEnd(ratio),
                   disk=(float(C.schwarzschild_radius(3.0, 0.5)), float(C.schwarzschild_radius(12.0, 0.5))))
    assert np.array_equal(status, o["status"]) and not outside.any()
    esc = status == 0
    pi_, di_ = C.schwarzschild_to_isotropic(o["exit_pos"][esc], o["exit_dir"][esc], 0.5)
    assert np.abs(end_loc[esc] * (ratio / r_obj) - pi_).max() / ratio < 1e-6 and np.abs(end_dir[esc] - di_).max() < 1e-6
    assert np.abs(np.linalg.norm(end_loc[esc], axis=1) - r_obj).max() < 1e-9
    hit = np.isfinite(o["disk_xy"][:, 0])
    assert np.array_equal(np.isfinite(dxy[:, 0]), hit) and hit.any()
    rr = np.linalg.norm(o["disk_xy"][hit], axis=1)
    assert np.abs(dxy[hit] - o["disk_xy"][hit] * (C.isotropic_radius(rr, 0.5) / rr)[:, None]).max() < 1e-6
    x, y, z, el, ed, mes = swi.ray_trace(d[5], locs[5], ratio_obj_to_blackhole=ratio)
    rr = np.sqrt(x * x + y * y + z * z)
    assert abs(rr[0] - ratio) < 1e-9 and abs(rr[-1] - ratio) < 1e-6 and (rr <= ratio + 1e-9).all()


def test_cam_call_shape():
    from blackhole_geodesic_calculator_b200 import adapters, raygen
    from oracle import port
    cam = adapters.RelativisticCamera(resolution=[16, 24], field_of_view=[0.5, 0.5], M=1.0).run()
    assert cam.ray_blackhole_hit.shape == (16, 24) and cam.ray_end.shape == (16, 24, 6)
    d = raygen.camera_rays(24, 16, 1, 0.5, 0.5, cam.rotation, jitter="none")
    p, hit = raygen.sphere_entry(cam.camera_location, d, 60.0)
    assert hit.all()
    o = port.trace(p, d)
    assert np.array_equal(cam.ray_blackhole_hit.reshape(-1), (o["status"] == 1).astype(np.int64))
    ok = o["status"] == 0
    assert np.abs(cam.ray_end.reshape(-1, 6)[ok, 3:6] - o["exit_dir"][ok]).max() < 1e-6   # CAM.py:228 reads [3:6]
    assert np.abs(cam.ray_end.reshape(-1, 6)[ok, 0:3] - o["exit_pos"][ok]).max() / 60.0 < 1e-6
    # a camera that partly misses the sphere: missing rays keep their flat direction and status -1
    cam2 = adapters.RelativisticCamera(resolution=[8, 8], field_of_view=[2.0, 2.0], M=1.0).run()
    assert (cam2.ray_status == 5).any() and (cam2.ray_status < 5).any()


def test_trace_sharded_nccl_two_gpus():
    """Interleaved sharding + NCCL gather, and the peer-memory frame (CUDA IPC over NVLink), on real devices
    (needs >= 2 GPUs; the gloo twin of the gather runs on CPU)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == ["none", "ok"]


def _nccl_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from blackhole_geodesic_calculator_b200 import api, distributed, raygen
        pos, d = raygen.config_bundle(64, 64, 1)
        tp = torch.from_numpy(pos).cuda()
        td = torch.from_numpy(d).cuda()
        out = distributed.trace_sharded(tp, td, dst=0)
        # same frame through the peer-memory route: both GPUs store into rank 0's HBM, no gather
        frame = distributed.PeerFrame(tp.shape[0], owner=0)
        try:
            peer = []
            # arrival flags by stream memory operations (default) and the NCCL fence, both routes, twice in a row so
            # that the second frame has to wait for the owner's release of the buffers
            for route, w, chunks, sync in (("stores", 0, 1, "flags"), ("courier", 64, 1, "flags"), ("copy", 0, 1, "flags"),
                                           ("copy", 64, 3, "nccl"), ("courier", 0, 1, "nccl"), ("auto", 64, 1, "flags")):
                if rank == 0:
                    frame.tensors()[0].fill_(float("nan"))
                    frame.tensors()[1].fill_(float("nan"))
                    frame.tensors()[2].fill_(-7)
                    torch.cuda.synchronize()
                dist.barrier()
                res = distributed.trace_sharded_peer(tp, td, frame, image_width=w, route=route, chunks=chunks, sync=sync)
                torch.cuda.synchronize()
                peer.append(None if res is None else [t.clone() for t in res])
                dist.barrier()
            torch.cuda.synchronize()
            # ragged frame: the last band is partial and 3 pieces do not divide the bands
            m = 3 * 8192 + 77
            tp2, td2 = tp.repeat(8, 1)[:m].contiguous(), td.repeat(8, 1)[:m].contiguous()
            frame2 = distributed.PeerFrame(m, owner=0)
            try:
                res2 = distributed.trace_sharded_peer(tp2, td2, frame2, route="copy", chunks=3)
                torch.cuda.synchronize()
                ok2a = True
                if rank == 0:
                    ok2a = all(torch.equal(a, b) for a, b in zip(res2, api.trace(tp2, td2)))
                    res2[2].fill_(-7)
                dist.barrier()
                res2 = distributed.trace_sharded_peer(tp2, td2, frame2, route="courier")
                torch.cuda.synchronize()
                ok2 = True
                if rank == 0:
                    ok2 = ok2a and all(torch.equal(a, b) for a, b in zip(res2, api.trace(tp2, td2)))
            finally:
                frame2.close()
            if rank == 0:
                ref = api.trace(tp, td)
                ok = all(torch.equal(a, b) for a, b in zip(out, ref))
                ok = ok and ok2 and all(torch.equal(a, b) for pr in peer for a, b in zip(pr, ref))
                q.put("ok" if ok else "mismatch")
            else:
                q.put("none" if out is None and peer == [None] * 6 else "unexpected")
        finally:
            frame.close()
    finally:
        dist.destroy_process_group()


def test_peer_frame_single_process_routes_equal_plain_trace():
    """PeerFrame without a process group (world 1): both delivery routes run through bhg_device_alloc, the `order`
    array / bhg_copy_rows on one GPU and must reproduce the plain call bit for bit, ragged frame included."""
    import torch
    from blackhole_geodesic_calculator_b200 import api, distributed, raygen
    pos, d = raygen.config_bundle(64, 64, 3)
    tp, td = torch.from_numpy(pos).cuda(), torch.from_numpy(d).cuda()
    for n, width in ((tp.shape[0], 64), (tp.shape[0], 0), (8192 + 77, 0)):
        p, q = tp[:n].contiguous(), td[:n].contiguous()
        ref = api.trace(p, q)
        frame = distributed.PeerFrame(n)
        try:
            for route, chunks in (("courier", 1), ("stores", 1), ("copy", 1), ("copy", 3)):
                # sentinels everywhere: a region a route failed to deliver must not pass on stale, identical data
                frame.tensors()[0].fill_(float("nan"))
                frame.tensors()[1].fill_(float("nan"))
                frame.tensors()[2].fill_(-7)
                got = distributed.trace_sharded_peer(p, q, frame, image_width=width, route=route, chunks=chunks)
                torch.cuda.synchronize()
                assert all(torch.equal(a, b) for a, b in zip(got, ref)), (n, width, route, chunks)
            # the optional orbital-plane mode through the courier (its finish path re-reads the entry state through the
            # band mapping of the in-place shard)
            ref_plane = api.trace(p, q, mode="plane")
            frame.tensors()[2].fill_(-7)
            got = distributed.trace_sharded_peer(p, q, frame, image_width=width, route="courier", mode="plane")
            torch.cuda.synchronize()
            assert all(torch.equal(a, b) for a, b in zip(got, ref_plane)), (n, width, "courier/plane")
        finally:
            frame.close()


def test_peer_frame_graph_replay_single_process():
    """The sharded-frame call is capturable in a CUDA graph (memset of the queue head, trace kernels on two streams,
    strided copy-engine deliveries) and the replay reproduces the plain call after the inputs changed in place."""
    import torch
    from blackhole_geodesic_calculator_b200 import api, distributed, raygen
    pos, d = raygen.config_bundle(64, 64, 2)
    tp, td = torch.from_numpy(pos).cuda(), torch.from_numpy(d).cuda()
    frame = distributed.PeerFrame(tp.shape[0])
    try:
        for route, chunks in (("stores", 1), ("copy", 2), ("courier", 1)):
            g = distributed.PeerFrameGraph(tp, td, frame, image_width=64, route=route, chunks=chunks)
            try:
                # new frame in the same static buffers: mirror the rays through the equatorial plane
                flip = torch.tensor([1.0, 1.0, -1.0], dtype=torch.float64, device="cuda")
                tp.mul_(flip)
                td.mul_(flip)
                got = g.replay()
                torch.cuda.synchronize()
                ref = api.trace(tp, td)
                assert all(torch.equal(a, b) for a, b in zip(got, ref)), route
            finally:
                g.close()
    finally:
        frame.close()


def test_trace_camera_f32_is_the_rounded_f64_result():
    """bhg_trace_camera_f32_host: FP64 integration, one rounding to float32 on store; 16 B/ray with want_pos=False.
    Also the buffer checks every entry point that takes caller arrays applies (a wrong dtype must not reach the DMA)."""
    import numpy as np
    from blackhole_geodesic_calculator_b200 import api, raygen
    cam = api.make_camera(raygen.CFG_CAMERA_POS, raygen.look_at_rotation(raygen.CFG_CAMERA_POS), 64, 48, 0.6, 0.6,
                          seed=7, jitter="philox")
    n = 64 * 48
    ep, ed, st = api.trace_camera(cam, n)
    fp, fd, fs = api.trace_camera_f32(cam, n, want_pos=True)
    assert np.array_equal(st, fs)
    assert np.array_equal(fd, ed.astype(np.float32)) and np.array_equal(fp, ep.astype(np.float32), equal_nan=True)
    none, fd2, fs2 = api.trace_camera_f32(cam, n)
    assert none is None and np.array_equal(fd2, fd) and np.array_equal(fs2, fs)
    import pytest
    with pytest.raises(ValueError):   # float64 arrays handed to the float32 entry point
        api.trace_camera_f32(cam, n, buffers=(None, np.empty((n, 3)), np.empty(n, np.int32)))
    with pytest.raises(ValueError):   # short status array
        api.trace_camera(cam, n, buffers=(np.empty((n, 3)), np.empty((n, 3)), np.empty(n - 1, np.int32)))
    with pytest.raises(ValueError):   # non-contiguous output
        api.trace_f32(np.zeros((8, 3), np.float32), np.ones((8, 3), np.float32),
                      out=(np.empty((8, 6), np.float32)[:, ::2], np.empty((8, 3), np.float32), np.empty(8, np.int32)))
    with pytest.raises(ValueError):   # disk annulus
        api.trace(np.array([[30.0, 0, 0]]), np.array([[-1.0, 0, 0]]), disk=(20.0, 6.0))


def test_camera_inside_the_sphere_starts_at_the_camera():
    """the RRE / CAM engines put the camera inside the curved region (RelativisticRenderEngineCamEdition.py:212,
    camera_location=[1e-5,-10,10] with r_sphere-style region around it): every ray then starts at the camera itself -
    device generator, host generator and the CAM adapter agree with the oracle on those rays."""
    import numpy as np
    from blackhole_geodesic_calculator_b200 import adapters, api, raygen
    from oracle import port
    cam_pos = (1e-5, -10.0, 10.0)
    rot = raygen.look_at_rotation(cam_pos)
    cam = api.make_camera(cam_pos, rot, 24, 16, 0.8, 0.8, seed=3, jitter="philox")
    n = 24 * 16
    pos, d, hit = api.generate_rays(cam, n, 60.0)
    pos, d = pos.cpu().numpy(), d.cpu().numpy()
    assert (hit.cpu().numpy() == 0).all() and np.allclose(pos, np.array(cam_pos)[None, :], rtol=0, atol=0)
    dirs = raygen.camera_rays(24, 16, 1, 0.8, 0.8, rot, 3, "philox")
    hp, hm = raygen.sphere_entry(cam_pos, dirs, 60.0)
    assert hm.all() and np.array_equal(hp, pos)
    ep, ed, st = api.trace_camera(cam, n)
    o = port.trace(pos, d)
    assert np.array_equal(st, o["status"]) and (st == 1).any() and (st == 0).any()
    esc = st == 0
    assert np.abs(ed[esc] - o["exit_dir"][esc]).max() < 1e-6 and np.abs(ep[esc] - o["exit_pos"][esc]).max() / 60.0 < 1e-6


def test_courier_route_as_the_very_first_call_of_a_process():
    """The courier kernel polls for bands the trace kernel completes.  As the FIRST CUDA work of a process nothing it
    depends on may still need loading (lazy module loading can wait for a context-wide synchronisation that a polling
    kernel never grants): every band must arrive, image-ordered, binned and ragged frames alike, with sentinels in the
    frame so that stale data cannot pass."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "courier_check.py")], capture_output=True, text=True,
                       timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("courier")]
    assert len(lines) == 3, r.stdout
    for l in lines:
        assert "mismatches [0, 0, 0]" in l and "nan left [0, 0]" in l and l.endswith("unwritten status 0"), l


def test_stream_flags_on_one_gpu():
    """bhg_stream_write32 / bhg_stream_wait_geq32 (driver stream memory operations resolved at run time): a stream
    that waits for a flag is released by a write issued later on another stream - the arrival / release protocol of
    the sharded frame, on local memory."""
    import torch
    from blackhole_geodesic_calculator_b200 import _lib
    lib = _lib.load()
    flag = torch.zeros(4, dtype=torch.int32, device="cuda")
    out = torch.zeros(1, dtype=torch.int32, device="cuda")
    a, b = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    _lib.check(lib.bhg_stream_wait_geq32(flag.data_ptr() + 4, 7, 0, a.cuda_stream))     # a: blocked until flag[1] >= 7
    with torch.cuda.stream(a):
        out.fill_(1)
    assert int(out.item()) == 0                                                         # still waiting (item() syncs the default stream only)
    _lib.check(lib.bhg_stream_write32(flag.data_ptr() + 4, 9, 0, b.cuda_stream))        # b: release
    a.synchronize()
    assert int(out.item()) == 1 and flag.tolist() == [0, 9, 0, 0]
    assert lib.bhg_stream_write32(flag.data_ptr() + 2, 1, 0, b.cuda_stream) != 0        # misaligned address is refused
