"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(include/bhgeo.h): the host entry point for numpy arrays, the device entry point for torch tensors.

Tolerances (BASELINE.json north_star): status bit-exact outside |b - 3 sqrt(3) M| <= 1e-2 M; exit position
within 1e-6 relative (to the sphere radius), exit direction within 1e-6 rad."""
import ctypes
import os

import numpy as np
import pytest

from conftest import (B_CRIT_BAND, COND_K, COND_SEEDS, COND_WELL, assert_conditioned_parity, assert_parity, golden_kwargs,
                      load_golden, ray_deviation)

pytestmark = pytest.mark.gpu

B_CRIT = 3.0 * np.sqrt(3.0)


@pytest.fixture(scope="module")
def api():
    from blackhole_geodesic_calculator_b200 import api as _api
    rc, errs = _api.selftest(0)
    print("selftest", rc, errs)
    assert rc == 0, errs
    return _api


def crit_band(pos, d, M):
    from blackhole_geodesic_calculator_b200 import raygen
    b = raygen.conserved_impact_parameter(pos, d, M)
    return np.abs(b - B_CRIT * M) <= B_CRIT_BAND * M


GOLDEN_FILES = ["cfg1_64x64.npz", "cfg5_nearcrit_3d.npz", "cfg5_nearcrit_plane.npz", "cfg3_sample.npz",
                "tight_16x16.npz", "rre_shape_16x16.npz", "maxstep_8x8.npz", "analytic_kat.npz"]


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_golden_parity_host_entry(api, name):
    """CUDA path vs the committed scipy golden vectors, numpy in / numpy out."""
    g = load_golden(name)
    kw = golden_kwargs(g)
    ep, ed, st, cnt = api.trace(g["entry_pos"], g["entry_dir"], return_counters=True, **kw)
    scale = kw["r_sphere"] if np.isfinite(kw["r_sphere"]) else 50.0
    band = crit_band(g["entry_pos"], g["entry_dir"], kw["M"])
    dpos, ddir = assert_parity(ep, ed, st, g["exit_pos"], g["exit_dir"], g["status"], scale, exclude=band)
    # same discrete algorithm: the step sequence itself must match (nfev = 2 + 6 attempts)
    integ = ~band & (g["status"] != 2)
    same = (2 + 6 * cnt[0][integ] == g["nfev"][integ]) & (cnt[1][integ] == g["n_accept"][integ])
    print(f"{name}: dpos {dpos:.2e} ddir {ddir:.2e} identical step sequences {same.mean() * 100:.2f}% "
          f"({(~same).sum()} differ)")
    assert same.mean() >= 0.999
    assert np.allclose(np.linalg.norm(ed[np.isin(st, (0, 1, 3))], axis=1), 1.0, atol=1e-12)


def test_edge_cases(api):
    g = load_golden("edge_cases.npz")
    kw = golden_kwargs(g)
    ep, ed, st = api.trace(g["entry_pos"], g["entry_dir"], **kw)
    assert np.array_equal(st, g["status"]), (st, g["status"])
    ok = np.isin(st, (0, 1, 3))
    assert np.abs(ep[ok] - g["exit_pos"][ok]).max() / 60.0 < 1e-6
    assert np.abs(ed[ok] - g["exit_dir"][ok]).max() < 1e-6
    assert np.isnan(ep[st == 2]).all() and np.isnan(ed[st == 2]).all()


def test_empty_and_ragged_sizes(api):
    ep, ed, st = api.trace(np.zeros((0, 3)), np.zeros((0, 3)))
    assert ep.shape == (0, 3) and st.shape == (0,)
    g = load_golden("cfg1_64x64.npz")
    from oracle import port
    for n in (1, 31, 33, 127, 129, 1000):
        ep, ed, st = api.trace(g["entry_pos"][:n], g["entry_dir"][:n])
        assert np.array_equal(st, g["status"][:n])
        assert np.abs(ep - g["exit_pos"][:n]).max() / 60.0 < 1e-6


def test_device_entry_layouts_order_and_thresholds_are_bitwise_identical(api):
    """Scheduling must not change arithmetic: SoA vs AoS, any refill threshold, any queue order."""
    import torch
    g = load_golden("cfg1_64x64.npz")
    n = g["entry_pos"].shape[0]
    dev = torch.device("cuda:0")
    pos = torch.from_numpy(g["entry_pos"]).to(dev)
    d = torch.from_numpy(g["entry_dir"]).to(dev)
    base = [t.cpu().numpy() for t in api.trace(pos, d)]
    assert_parity(base[0], base[1], base[2], g["exit_pos"], g["exit_dir"], g["status"], 60.0,
                  exclude=crit_band(g["entry_pos"], g["entry_dir"], 1.0))
    for T in (1, 4, 12, 32):
        out = [t.cpu().numpy() for t in api.trace(pos, d, refill_threshold=T)]
        for a, b in zip(base, out):
            assert np.array_equal(a, b, equal_nan=True), f"threshold {T}"
    for wdt in (64, 8, 24, 100):  # 100: not a multiple of 8 -> hint ignored
        out = [t.cpu().numpy() for t in api.trace(pos, d, image_width=wdt)]
        for a, b in zip(base, out):
            assert np.array_equal(a, b, equal_nan=True), f"image_width {wdt}"
    # SoA planes + a random queue order
    soa_in = torch.cat([pos.t().contiguous(), d.t().contiguous()]).contiguous()  # [6, n]
    soa_out = torch.empty_like(soa_in)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    order = torch.randperm(n, device=dev, dtype=torch.int32)
    params = api.make_params()
    api.trace_device(soa_in.data_ptr(), None, soa_out.data_ptr(), None, status.data_ptr(), None, order.data_ptr(),
                     n, api.LAYOUT_SOA, params, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(soa_out[:3].t().cpu().numpy(), base[0], equal_nan=True)
    assert np.array_equal(soa_out[3:].t().cpu().numpy(), base[1], equal_nan=True)
    assert np.array_equal(status.cpu().numpy(), base[2])


def test_pinned_host_buffers(api):
    g = load_golden("cfg3_sample.npz")
    n = g["entry_pos"].shape[0]
    pos = api.pinned_empty((n, 3))
    d = api.pinned_empty((n, 3))
    pos[:] = g["entry_pos"]
    d[:] = g["entry_dir"]
    ep, ed, st = api.trace(pos, d)
    assert_parity(ep, ed, st, g["exit_pos"], g["exit_dir"], g["status"], 60.0)
    api.pinned_free(pos)
    api.pinned_free(d)


def test_plane_mode_against_its_cpu_restatement_and_analytic(api):
    from oracle import port
    g = load_golden("cfg5_nearcrit_3d.npz")
    ep, ed, st, cnt = api.trace(g["entry_pos"], g["entry_dir"], mode="plane", return_counters=True)
    o = port.trace(g["entry_pos"], g["entry_dir"], mode=1)
    band = np.abs(g["b"] - B_CRIT) <= B_CRIT_BAND
    assert_parity(ep, ed, st, o["exit_pos"], o["exit_dir"], o["status"], 60.0, exclude=band)
    assert (cnt[0][~band] == o["n_attempt"][~band]).mean() > 0.999
    # analytic deflection at tight tolerance
    k = load_golden("analytic_kat.npz")
    ep, ed, st = api.trace(k["entry_pos"], k["entry_dir"], rtol=1e-12, atol=1e-14, mode="plane")
    e_in = np.arctan2(k["entry_pos"][:, 1], k["entry_pos"][:, 0])
    e_out = np.arctan2(ep[:, 1], ep[:, 0])
    swept = np.mod(e_in - e_out, 2 * np.pi)
    kk = np.round((k["dphi_analytic"] - swept) / (2 * np.pi))
    assert np.abs(swept + 2 * np.pi * kk - k["dphi_analytic"]).max() < 5e-8


def test_analytic_deflection_parity_mode(api):
    k = load_golden("analytic_kat.npz")
    ep, ed, st = api.trace(k["entry_pos"], k["entry_dir"], rtol=1e-12, atol=1e-14)
    assert (st == 0).all()
    e_in = np.arctan2(k["entry_pos"][:, 1], k["entry_pos"][:, 0])
    e_out = np.arctan2(ep[:, 1], ep[:, 0])
    swept = np.mod(e_in - e_out, 2 * np.pi)
    kk = np.round((k["dphi_analytic"] - swept) / (2 * np.pi))
    assert np.abs(swept + 2 * np.pi * kk - k["dphi_analytic"]).max() < 5e-8


def test_full_frame_properties_and_subsample_parity(api):
    """BASELINE config 2 (1024 x 1024 x 5 spp = 5 242 880 rays) at full size: size-independent properties on
    every ray, and per-ray parity against the oracle's C restatement on EVERY ray."""
    import torch
    from blackhole_geodesic_calculator_b200 import raygen
    from oracle import port
    pos, d = raygen.config_bundle(1024, 1024, 5, jitter="philox")
    n = pos.shape[0]
    assert n == 5242880
    dev = torch.device("cuda:0")
    tp, td = torch.from_numpy(pos).to(dev), torch.from_numpy(d).to(dev)
    ep, ed, st, cnt = api.trace(tp, td, return_counters=True)
    torch.cuda.synchronize()
    att, acc, integ = api.sum_counters(cnt.data_ptr(), st.data_ptr(), n)
    ep, ed, st, cnt = ep.cpu().numpy(), ed.cpu().numpy(), st.cpu().numpy(), cnt.cpu().numpy()
    assert att == int(cnt[0].sum()) and acc == int(cnt[1].sum()) and integ == n
    assert set(np.unique(st)) <= {0, 1}
    esc = st == 0
    assert np.abs(np.linalg.norm(ep[esc], axis=1) - 60.0).max() < 1e-9        # exit on the sphere
    assert np.abs(np.linalg.norm(ep[~esc], axis=1) - 2.01).max() < 1e-9       # capture on r_s + eps
    assert np.abs(np.linalg.norm(ed, axis=1) - 1.0).max() < 1e-12            # unit directions
    assert (np.sum(ep[esc] * ed[esc], axis=1) > 0).all()                      # leaving outward
    # the shadow: captured iff b < b_crit.  This is physics, not parity: at rtol=1e-3 the reference algorithm
    # itself misclassifies rays up to |b - b_c| = 0.13 M in 3-D spherical coordinates (measured with the
    # oracle on this frame), so the physical check uses a 0.2 M margin; parity is checked below.
    b = raygen.conserved_impact_parameter(pos, d, 1.0)
    away = np.abs(b - B_CRIT) > B_CRIT_BAND
    far = np.abs(b - B_CRIT) > 0.2
    assert np.array_equal(st[far] == 1, b[far] < B_CRIT)
    # conserved angular momentum direction: exit state stays in the entry orbital plane
    nrm = np.cross(pos, d)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    big = b > 1.0
    assert np.abs(np.sum(nrm * ep, axis=1))[big].max() / 60.0 < 0.15   # default-tolerance drift only (oracle: 0.093)
    # every ray against the oracle's C restatement under the per-ray conditioning rule (conftest.py): rays whose
    # oracle result does not move under 1-ulp RHS jitter (all but a handful: the pole-grazing ones) must have
    # identical step counts and agree to 1e-6 without exception; statuses equal on every ray, also inside the
    # +-1e-2 M band around b_crit.
    o = port.trace(pos, d)
    cond, _ = port.conditioning(pos, d, base=o, seeds=COND_SEEDS)
    res = assert_conditioned_parity(ep, ed, st, cnt[0], cnt[1], o, cond, 60.0, label="config 2 full frame")
    assert res["ill_conditioned"] < 1e-4 * n and res["steps_differ"] <= 5
    print(f"full frame: attempts/ray {att / n:.2f}, captured {100 * (~esc).mean():.3f}%, in the +-1e-2 M band "
          f"{int((~away).sum())} rays (statuses equal there too)")


def test_time_reversal_on_gpu(api):
    g = load_golden("cfg3_sample.npz")
    ep, ed, st = api.trace(g["entry_pos"], g["entry_dir"], rtol=1e-11, atol=1e-13)
    esc = st == 0
    bp, bd, bs = api.trace(ep[esc] * (1 - 1e-12), -ed[esc], rtol=1e-11, atol=1e-13)
    assert (bs == 0).all()
    assert np.abs(bp - g["entry_pos"][esc]).max() / 60.0 < 1e-6
    assert np.abs(bd + g["entry_dir"][esc]).max() < 1e-6


def test_error_codes_on_gpu(api):
    from blackhole_geodesic_calculator_b200 import _lib
    lib = _lib.load()
    p = api.make_params()
    rc = lib.bhg_trace_schwarzschild_f64(None, None, None, None, None, None, None, 8, 1, ctypes.byref(p), 0, None)
    assert rc == -1 and b"NULL" in lib.bhg_last_error_string()
    rc = lib.bhg_trace_schwarzschild_f64(None, None, None, None, None, None, None, 0, 7, ctypes.byref(p), 0, None)
    assert rc == -1
    rc = lib.bhg_trace_schwarzschild_f64(None, None, None, None, None, None, None, 0, 1, ctypes.byref(p), 63, None)
    assert rc == -3


def test_device_raygen_matches_host_generator(api):
    """device-side primary rays vs the host generator (same Philox stream, same order)."""
    from blackhole_geodesic_calculator_b200 import raygen
    rot = raygen.look_at_rotation(raygen.CFG_CAMERA_POS)
    for (w, h, spp, jit) in ((64, 48, 2, "philox"), (40, 24, 1, "none")):
        n = w * h * spp
        cam = api.make_camera(raygen.CFG_CAMERA_POS, rot, w, h, 0.6, 0.6, seed=42, jitter=jit)
        pos, d, hit = (t.cpu().numpy() for t in api.generate_rays(cam, n, 60.0))
        hd = raygen.camera_rays(w, h, spp, 0.6, 0.6, rot, 42, jit)
        hp, hhit = raygen.sphere_entry(raygen.CFG_CAMERA_POS, hd, 60.0)
        assert np.array_equal(hit == 0, hhit)
        assert np.abs(d - hd).max() < 4e-16                      # a few ulp (matmul summation order)
        assert np.abs(pos[hhit] - hp[hhit]).max() / 60.0 < 1e-14
    # a wide camera: some rays miss the sphere
    cam = api.make_camera(raygen.CFG_CAMERA_POS, rot, 32, 32, 2.0, 2.0, jitter="none")
    pos, d, hit = (t.cpu().numpy() for t in api.generate_rays(cam, 1024, 60.0))
    assert (hit == 5).any() and (hit == 0).any() and np.isnan(pos[hit == 5]).all()
    # first_ray offsets address the same stream
    cam2 = api.make_camera(raygen.CFG_CAMERA_POS, rot, 32, 32, 2.0, 2.0, jitter="none", first_ray=500)
    pos2, d2, hit2 = (t.cpu().numpy() for t in api.generate_rays(cam2, 524, 60.0))
    assert np.array_equal(d2, d[500:]) and np.array_equal(hit2, hit[500:])


def test_fused_camera_trace_equals_generate_then_trace(api):
    """The fused kernel must give bit-identical results to generate_rays + trace, on host and device outputs,
    with and without exit positions, including rays that miss the sphere."""
    import torch
    from blackhole_geodesic_calculator_b200 import raygen
    from oracle import port
    rot = raygen.look_at_rotation(raygen.CFG_CAMERA_POS)
    for fov, w, h in ((0.6, 64, 64), (1.2, 48, 40)):
        n = w * h
        cam = api.make_camera(raygen.CFG_CAMERA_POS, rot, w, h, fov, fov, seed=7, jitter="philox")
        pos, d, hit = api.generate_rays(cam, n, 60.0)
        ok = (hit == 0).cpu().numpy()
        ep0, ed0, st0 = (t.cpu().numpy() for t in api.trace(pos[hit == 0].contiguous(), d[hit == 0].contiguous()))
        ep, ed, st, cnt = api.trace_camera(cam, n, return_counters=True)
        assert np.array_equal(st[ok], st0) and (st[~ok] == 5).all()
        assert np.array_equal(ep[ok], ep0) and np.array_equal(ed[ok], ed0)
        assert np.isnan(ep[~ok]).all() and np.array_equal(ed[~ok], d.cpu().numpy()[~ok])
        # directions only, torch outputs
        _, ed_t, st_t = api.trace_camera(cam, n, want_pos=False, out="torch")
        torch.cuda.synchronize()
        assert np.array_equal(ed_t.cpu().numpy(), ed) and np.array_equal(st_t.cpu().numpy(), st)
        # and against the oracle on the device-generated rays
        o = port.trace(pos.cpu().numpy()[ok], d.cpu().numpy()[ok])
        band = crit_band(pos.cpu().numpy()[ok], d.cpu().numpy()[ok], 1.0)
        assert_parity(ep[ok], ed[ok], st[ok], o["exit_pos"], o["exit_dir"], o["status"], 60.0, exclude=band)
        assert (cnt[0][ok][~band] == o["n_attempt"][~band]).mean() > 0.999


def test_sky_uv_mapping(api):
    """the equirectangular lookup coordinates of background_hit (RRE.py:366-378)."""
    import torch
    from blackhole_geodesic_calculator_b200 import raygen
    g = load_golden("cfg1_64x64.npz")
    d = torch.from_numpy(g["exit_dir"]).cuda()
    st = torch.from_numpy(g["status"].astype(np.int32)).cuda()
    uv = api.sky_uv(d, st).cpu().numpy()
    ok = g["status"] == 0
    theta = 1 - np.arccos(g["exit_dir"][:, 2]) / np.pi          # the reference's expressions, literally
    phi = np.arctan2(g["exit_dir"][:, 1], g["exit_dir"][:, 0]) / np.pi
    ref = np.stack([-phi, 2 * theta - 1], axis=1)
    assert np.abs(uv[ok] - ref[ok]).max() < 1.2e-7            # float32 storage of an FP64 result
    assert np.isnan(uv[~ok]).all()
    # fused camera -> uv on the host
    rot = raygen.look_at_rotation(raygen.CFG_CAMERA_POS)
    cam = api.make_camera(raygen.CFG_CAMERA_POS, rot, 64, 64, 0.6, 0.6, seed=3, jitter="philox")
    uv_h, st_h = api.trace_camera_sky(cam, 4096)
    _, ed, st2 = api.trace_camera(cam, 4096, want_pos=False)
    assert np.array_equal(st_h, st2)
    esc = st2 == 0
    ref = np.stack([-np.arctan2(ed[:, 1], ed[:, 0]) / np.pi, 2 * (1 - np.arccos(ed[:, 2]) / np.pi) - 1], axis=1)
    assert np.abs(uv_h[esc] - ref[esc]).max() < 1.2e-7 and np.isnan(uv_h[st2 == 1]).all()


def test_disk_crossing_event(api):
    """first equatorial-plane crossing inside the annulus, located in flight on the dense output;
    CUDA vs the scipy golden (non-terminal solve_ivp event) and, on a larger bundle, vs the C port."""
    import torch
    g = load_golden("disk_crossing.npz")
    kw = golden_kwargs(g)
    disk = (float(g["disk_r_in"]), float(g["disk_r_out"]))
    ep, ed, st, dxy = api.trace(g["entry_pos"], g["entry_dir"], disk=disk, **kw)
    band = crit_band(g["entry_pos"], g["entry_dir"], 1.0)
    assert_parity(ep, ed, st, g["exit_pos"], g["exit_dir"], g["status"], 60.0, exclude=band)
    hit = np.isfinite(g["disk_xy"][:, 0])
    assert np.array_equal(np.isfinite(dxy[:, 0])[~band], hit[~band])
    m = hit & ~band
    assert np.abs(dxy[m] - g["disk_xy"][m]).max() / 20.0 < 1e-6
    # the event never changes the integration itself
    ep0, ed0, st0 = api.trace(g["entry_pos"], g["entry_dir"], **kw)
    assert np.array_equal(ep, ep0, equal_nan=True) and np.array_equal(ed, ed0, equal_nan=True) and np.array_equal(st, st0)
    # device entry, larger bundle, C port as checker
    from blackhole_geodesic_calculator_b200 import raygen
    from oracle import port
    pos, d = raygen.config_bundle(256, 256, 1, jitter="philox")
    out = api.trace(torch.from_numpy(pos).cuda(), torch.from_numpy(d).cuda(), disk=disk, image_width=256)
    dxy_t = out[-1].cpu().numpy()
    o = port.trace(pos, d, disk=disk)
    band = crit_band(pos, d, 1.0)
    hit = np.isfinite(o["disk_xy"][:, 0])
    assert (np.isfinite(dxy_t[:, 0]) == hit)[~band].all() and hit.sum() > 1000
    m = hit & ~band
    assert np.abs(dxy_t[m] - o["disk_xy"][m]).max() / 20.0 < 1e-6
    with pytest.raises(Exception):
        api.trace(g["entry_pos"], g["entry_dir"], disk=disk, mode="plane")


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_randomised_parameters_against_c_port(api, seed):
    """Random mass / sphere radius / tolerances / max_step / lambda_max / horizon offset, camera bundles from random
    directions (inside and outside the sphere): CUDA vs the C restatement, both modes."""
    from blackhole_geodesic_calculator_b200 import raygen
    from oracle import port
    rng = np.random.default_rng(seed)
    M = float(rng.uniform(0.3, 3.0))
    R = float(rng.uniform(12.0, 90.0)) * M
    rtol = float(10.0 ** rng.uniform(-8, -3))
    atol = float(rtol * 10.0 ** rng.uniform(-4, -2))
    max_step = float(rng.choice([np.inf, np.inf, 5.0 * M, 20.0 * M]))
    eps = float(rng.choice([0.01, 0.05, 0.001])) * M
    lam = float(rng.choice([0.0, 0.0, 3.0 * R]))  # 0 -> default 10 R
    cam = rng.normal(size=3)
    cam *= rng.uniform(1.3, 3.0) * R / np.linalg.norm(cam)
    half = math_asin(R / np.linalg.norm(cam))
    rot = raygen.look_at_rotation(tuple(cam))
    d = raygen.camera_rays(48, 32, 1, 1.2 * half, 1.2 * half, rot, seed, "philox")
    pos, hit = raygen.sphere_entry(cam, d, R)
    pos, d = pos[hit], d[hit]
    kw = dict(M=M, r_sphere=R, rtol=rtol, atol=atol, max_step=max_step, eps_horizon=eps,
              lambda_max=None if lam == 0.0 else lam)
    for mode, pmode in (("parity", 0), ("plane", 1)):
        ep, ed, st, cnt = api.trace(pos, d, mode=mode, return_counters=True, **kw)
        o = port.trace(pos, d, mode=pmode, **kw)
        b = raygen.conserved_impact_parameter(pos, d, M)
        band = np.abs(b - B_CRIT * M) <= B_CRIT_BAND * M
        assert_parity(ep, ed, st, o["exit_pos"], o["exit_dir"], o["status"], R, exclude=band)
        same = cnt[0][~band] == o["n_attempt"][~band]
        assert same.mean() > 0.995, (mode, same.mean())


def math_asin(x):
    import math
    return math.asin(x)


def test_pageable_and_pinned_host_paths_agree(api):
    """Plain numpy arrays go through the pinned bounce pipeline; pinned arrays are copied directly: same bits."""
    from blackhole_geodesic_calculator_b200 import raygen
    pos, d = raygen.config_bundle(512, 384, 3, jitter="philox")   # 589 824 rays: several pipeline chunks, ragged tail
    pos, d = pos[:-77], d[:-77]
    n = pos.shape[0]
    a = api.trace(pos, d, return_counters=True, disk=(6.0, 20.0))
    ppos, pd = api.pinned_empty((n, 3)), api.pinned_empty((n, 3))
    ppos[:] = pos
    pd[:] = d
    # pinned inputs alone still take the bounce path (outputs are pageable); all-pinned is exercised by bench.py
    b = api.trace(ppos, pd, return_counters=True, disk=(6.0, 20.0))
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    import torch
    c = api.trace(torch.from_numpy(pos).cuda(), torch.from_numpy(d).cuda(), return_counters=True, disk=(6.0, 20.0))
    for x, y in zip(a, c):
        assert np.array_equal(x, y.cpu().numpy(), equal_nan=True)
    api.pinned_free(ppos)
    api.pinned_free(pd)


def test_float32_io_variant(api):
    """float32 in / out (Blender's native precision): same FP64 integration of the exactly widened inputs."""
    g = load_golden("cfg1_64x64.npz")
    p32, d32 = g["entry_pos"].astype(np.float32), g["entry_dir"].astype(np.float32)
    ep, ed, st = api.trace_f32(p32, d32)
    ep64, ed64, st64 = api.trace(p32.astype(np.float64), d32.astype(np.float64))
    assert ep.dtype == np.float32 and np.array_equal(st, st64)
    assert np.array_equal(ep, ep64.astype(np.float32), equal_nan=True)
    assert np.array_equal(ed, ed64.astype(np.float32), equal_nan=True)
    # and it is within float32 granularity of the float64 golden answer for the typical ray (the 1e-7 input
    # rounding is amplified on near-critical rays, so only the bulk of the distribution is bounded)
    ok = (g["status"] == 0) & (st == 0)
    assert ok.mean() > 0.98
    dev = np.abs(ep[ok] - g["exit_pos"][ok]).max(axis=1) / 60.0
    assert np.median(dev) < 1e-6 and np.percentile(dev, 90) < 1e-4


def test_flat_limit_and_concurrent_callers(api):
    """M -> 0 gives straight chords; two host threads may call the library at the same time (ctypes drops the GIL)."""
    import threading
    from blackhole_geodesic_calculator_b200 import raygen
    pos, d = raygen.config_bundle(64, 64, 1)
    ep, ed, st = api.trace(pos, d, M=1e-9, rtol=1e-10, atol=1e-12)
    chord = -2 * np.sum(pos * d, axis=1)
    assert (st == 0).all()
    assert np.abs(ep - (pos + chord[:, None] * d)).max() < 1e-5 and np.abs(ed - d).max() < 1e-7
    g = load_golden("cfg1_64x64.npz")
    results = [None, None]

    def work(i):
        for _ in range(5):
            results[i] = api.trace(g["entry_pos"], g["entry_dir"])

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for r in results:
        assert np.array_equal(r[2], g["status"])
        assert np.array_equal(r[0], results[0][0], equal_nan=True)


@pytest.mark.parametrize("name", ["polyline_rre.npz", "polyline_sphere.npz"])
def test_polyline_samples(api, name):
    """SURVEY 8f row 2 (second half): positions on linspace(0, curve_end, K) up to termination, as solve_ivp's t_eval
    returns them to the reference (RelativisticRenderEngine.py:293-294,299-300)."""
    g = load_golden(name)
    kw = golden_kwargs(g)
    K = int(g["polyline"])
    ep, ed, st, poly, cnt = api.trace(g["entry_pos"], g["entry_dir"], polyline=K, **kw)
    assert np.array_equal(st, g["status"])
    assert np.array_equal(cnt, g["poly_count"])
    scale = 60.0 if np.isfinite(kw["r_sphere"]) else 50.0
    for i in range(len(cnt)):
        c = cnt[i]
        assert np.abs(poly[i, :c] - g["poly_xyz"][i, :c]).max() / scale < 1e-6
        assert np.isnan(poly[i, c:]).all()
    # sample 0 is the entry point; without the option nothing else changes
    assert np.abs(poly[:, 0] - g["entry_pos"]).max() < 1e-12
    ep0, ed0, st0 = api.trace(g["entry_pos"], g["entry_dir"], **kw)
    assert np.array_equal(ep0, ep, equal_nan=True) and np.array_equal(ed0, ed, equal_nan=True)
    # together with the disk event
    if np.isfinite(kw["r_sphere"]):
        out = api.trace(g["entry_pos"], g["entry_dir"], polyline=K, disk=(6.0, 20.0), **kw)
        assert np.array_equal(out[4], poly, equal_nan=True) and out[3].shape == (len(cnt), 2)


def test_full_size_configs_3_and_5_against_c_port(api):
    """BASELINE configs 3 (1920x1080 frame, 2.03 M rays) and 5 (2^20 near-critical rays, in the equatorial plane and
    in random planes) at full size against the C restatement, every ray, under the per-ray conditioning rule
    (conftest.py; adjudicated three-way against real scipy in profiles/r2a_adjudication.json):

    * statuses equal on every ray of all three sets, inside and outside the +-1e-2 M band around b_crit;
    * every ray whose oracle result is insensitive to 1-ulp RHS jitter (conditioning < 1e-7: all of the in-plane set,
      all but a few pole-grazing rays of config 3, 99.6 % of config 5 in random planes): identical step counts and
      exit states within 1e-6, no exceptions (in-plane set: 1e-7);
    * the ill-conditioned rest (near-critical rays that cross the polar region of the reference's coordinates on
      every half orbit): within max(1e-6, 10 x conditioning), at most 1 ray in 10^5 beyond that.  scipy and its own C
      restatement differ from each other on the same rays by the same amounts (79 / 624 rays with different step
      counts / beyond 1e-6 there, 91 / 678 for the CUDA path)."""
    from blackhole_geodesic_calculator_b200 import raygen
    from oracle import port

    def compare(p, d, label):
        p, d = np.ascontiguousarray(p), np.ascontiguousarray(d)
        ep, ed, st, cnt = api.trace(p, d, return_counters=True)
        o = port.trace(p, d)
        cond, _ = port.conditioning(p, d, base=o, seeds=COND_SEEDS)
        res = assert_conditioned_parity(ep, ed, st, cnt[0], cnt[1], o, cond, 60.0, label=label)
        return res, ray_deviation(ep, ed, o["exit_pos"], o["exit_dir"], 60.0), st

    res, dev, st = compare(*raygen.random_impact_bundle(None), "config 3")
    assert res["steps_differ"] == 0 and res["ill_conditioned"] < 200 and dev[st == 0].max() < 1e-3

    p5, d5, _ = raygen.near_critical_bundle(1 << 20, in_plane=True)
    res, dev, st = compare(p5, d5, "config 5 in-plane")
    assert res["steps_differ"] == 0 and res["ill_conditioned"] == 0 and dev[st == 0].max() < 1e-7

    p5, d5, _ = raygen.near_critical_bundle(1 << 20, in_plane=False)
    res, dev, st = compare(p5, d5, "config 5 random planes")
    assert res["steps_differ"] < 200 and res["ill_conditioned"] < 0.01 * len(st)


def test_adjudicated_outliers_against_scipy_golden(api):
    """The rays on which GPU, scipy and the C restatement disagreed pairwise in the round-2 adjudication
    (tests/golden/parity_outliers.npz: config 5 random planes - every outlier of the 2^20 rays; configs 3 and 2 - the
    GPU<->port outliers; plus control rays), CUDA path against the REAL scipy results, same per-ray rule."""
    g = load_golden("parity_outliers.npz")
    for name in ("cfg5", "cfg3", "cfg2"):
        p, d = g[name + "_entry_pos"], g[name + "_entry_dir"]
        ep, ed, st, cnt = api.trace(p, d, return_counters=True)
        ref = dict(exit_pos=g[name + "_scipy_pos"], exit_dir=g[name + "_scipy_dir"], status=g[name + "_scipy_status"],
                   n_attempt=g[name + "_scipy_attempt"], n_accept=g[name + "_scipy_accept"])
        # the set is a selection of the worst rays of 1e6: the 1-in-1e5 allowance is taken over the full set size
        dev = ray_deviation(ep, ed, ref["exit_pos"], ref["exit_dir"], 60.0)
        cond = g[name + "_conditioning"]
        assert np.array_equal(st, ref["status"]), name
        cmp_ = np.isin(ref["status"], (0, 3))
        well = cond < COND_WELL
        same = (cnt[0] == ref["n_attempt"]) & (cnt[1] == ref["n_accept"])
        assert same[well].all() and (dev[cmp_ & well] <= 1e-6).all(), name
        viol = cmp_ & ~well & (dev > np.maximum(1e-6, COND_K * cond))
        print(f"{name}: {len(st)} rays, ill-conditioned {int((cmp_ & ~well).sum())}, steps differ {int((~same).sum())}, "
              f"beyond 1e-6 {int((cmp_ & (dev > 1e-6)).sum())}, beyond the conditioned bound {int(viol.sum())}")
        assert viol.sum() <= 10, (name, np.nonzero(viol)[0], dev[viol], cond[viol])


def test_queue_order_features_never_change_a_result():
    """Cost binning (BHG_BIN), long-rays-first (BHG_HOT) and the pre-pass (BHG_PREP) reorder or relocate work; every ray's
    arithmetic is its own, so the outputs must be bit-identical whatever is switched on - checked on an unordered
    near-critical bundle (goes through the binning) and an image-ordered one small enough for the long-ray list."""
    import hashlib
    import subprocess
    import sys
    code = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
from blackhole_geodesic_calculator_b200 import api, raygen
h = hashlib.sha256()
p5, d5, _ = raygen.near_critical_bundle(1 << 16, in_plane=False)
for a in api.trace(p5, d5, return_counters=True): h.update(np.ascontiguousarray(a).tobytes())
p2, d2 = raygen.config_bundle(256, 256, 2, jitter="philox")
for a in api.trace(p2, d2, image_width=256, return_counters=True): h.update(np.ascontiguousarray(a).tobytes())
for a in api.trace(p2, d2, return_counters=True): h.update(np.ascontiguousarray(a).tobytes())
print(h.hexdigest())
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = {}
    for env in ({}, {"BHG_BIN": "0"}, {"BHG_BIN": "2"}, {"BHG_HOT": "0"}, {"BHG_PREP": "0"}):
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        digests[str(env)] = r.stdout.strip().splitlines()[-1]
    assert len(set(digests.values())) == 1, digests
