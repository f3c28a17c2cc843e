"""`trace()` — the batched drop-in for the reference's per-ray curvedpy call.

Replaces, for a whole frame or tile at once,
  k_xyz, x_xyz, result = self.GeoInt.calc_trajectory(k0_xyz, x0_xyz, max_step=..., curve_end=..., ...)
      (/root/reference/raytracer/RelativisticRenderEngine.py:293-294, end state :307-308, status :296-297)
  x, y, z, end_loc, end_dir, mes = self.SW.ray_trace(direction, loc_hit=loc, ...)
      (/root/reference/raytracer/LimitedRelativisticRenderEngine.py:273-278, status :308-314)
with
  exit_pos, exit_dir, status = trace(entry_pos[N,3], entry_dir[N,3], M, r_sphere, rtol, atol)

numpy in / numpy out goes through the host C entry point; torch CUDA tensors in / out stay on the device
(torch is optional and only imported when a tensor is passed).
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np

from . import _lib
from ._lib import BhgCamera, BhgExtras, BhgParams

ESCAPED, CAPTURED, START_INSIDE_HOLE, LAMBDA_EXHAUSTED, STEP_FAILED, MISSED_SPHERE = 0, 1, 2, 3, 4, 5
STATUS_NAMES = {0: "ESCAPED", 1: "CAPTURED", 2: "START_INSIDE_HOLE", 3: "LAMBDA_EXHAUSTED", 4: "STEP_FAILED",
                5: "MISSED_SPHERE"}
MODES = {"parity": 0, "plane": 1}
COORDS = {"schwarzschild": 0, "isotropic": 1}
LAYOUT_SOA, LAYOUT_AOS = 0, 1


def make_params(M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, max_step=math.inf, eps_horizon=0.01,
                lambda_max=None, mode="parity", refill_threshold=0, image_width=0, coords="schwarzschild") -> BhgParams:
    if mode not in MODES:
        raise ValueError(f"mode must be one of {sorted(MODES)}, got {mode!r}")
    if coords not in COORDS:
        raise ValueError(f"coords must be one of {sorted(COORDS)}, got {coords!r}")
    if max_step is None or max_step == -1:  # the reference maps -1 to inf (RelativisticRenderEngine.py:59-60)
        max_step = math.inf
    return BhgParams(float(M), float(r_sphere), float(rtol), float(atol), float(max_step), float(eps_horizon),
                     0.0 if lambda_max is None else float(lambda_max), MODES[mode], int(refill_threshold),
                     int(image_width), COORDS[coords])


def _check_out(a, shape, dtype, name):
    """Caller-supplied output buffer: the C ABI receives a raw pointer, so shape, dtype and layout are checked here."""
    if not isinstance(a, np.ndarray) or a.shape != tuple(shape) or a.dtype != np.dtype(dtype) or not a.flags.c_contiguous \
            or not a.flags.writeable:
        raise ValueError(f"{name} must be a writeable C-contiguous numpy array of shape {tuple(shape)} and dtype "
                         f"{np.dtype(dtype).name}; got {type(a).__name__}"
                         + (f" {a.shape} {a.dtype}" if isinstance(a, np.ndarray) else ""))


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def trace(entry_pos, entry_dir, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, *, max_step=math.inf,
          eps_horizon=0.01, lambda_max=None, mode="parity", refill_threshold=0, image_width=0, device=0,
          return_counters=False, disk=None, out=None, polyline=None, coords="schwarzschild"):
    """Integrate N Schwarzschild null geodesics from sphere entry to exit or capture.

    coords : chart of every position / direction / radius passed in and returned: "schwarzschild" (default; the
        spherical chart of README.md:162-172 read as x = r sin th cos ph ...) or "isotropic" (the Cartesian chart of the
        reference's older `SchwarzschildGeodesic` generation, the one README Fig. 5 / 6 are drawn in; include/bhgeo.h).
    entry_pos, entry_dir : [N,3] float64, BH-centred position and coordinate direction (numpy arrays on the
        host, or torch CUDA tensors, which are processed in place on their device and stream).
    image_width : optional scheduling hint — the rays are a row-major image of this width (the reference's
        s -> y -> x order); warps then integrate 8 x 4 pixel tiles.  Never changes results.
    out : optional (exit_pos, exit_dir, status) numpy arrays to fill (reuse them across frames — ideally
        `pinned_empty` ones: freshly allocated pageable outputs cost more in page faults than the trace itself).
    disk : optional (r_in, r_out): also return disk_xy[N,2], the first crossing of the equatorial plane z = 0
        with r_in <= r <= r_out (NaN = none) — the in-flight form of the reference's checkHitDisk
        (LimitedRelativisticRenderEngine.py:413-438).  Parity mode only.
    polyline : optional K >= 2 (needs an explicit lambda_max): also return (poly_xyz[N,K,3], poly_count[N]), the
        positions at lambda = linspace(0, lambda_max, K) up to each ray's termination (NaN beyond) - what curvedpy
        returns for nr_points_curve (RelativisticRenderEngine.py:293-294).  Parity mode, numpy inputs.
    Returns (exit_pos[N,3], exit_dir[N,3] unit-norm, status[N] int32) and, with return_counters, an
    int32 [2,N] array of (RK45 attempts, accepted steps), then disk_xy, then (poly_xyz, poly_count) if requested.
    """
    params = make_params(M, r_sphere, rtol, atol, max_step, eps_horizon, lambda_max, mode, refill_threshold,
                         image_width, coords)
    if disk is not None:
        r_in, r_out = float(disk[0]), float(disk[1])
        if not (0.0 <= r_in <= r_out and r_out > 0.0 and math.isfinite(r_out)):
            raise ValueError(f"disk=(r_in, r_out) needs 0 <= r_in <= r_out, r_out > 0 and finite; got {disk}")
    if _is_torch(entry_pos):
        if polyline is not None:
            raise ValueError("polyline output is available for numpy inputs")
        return _trace_torch(entry_pos, entry_dir, params, return_counters, disk)
    lib = _lib.load()
    pos = np.ascontiguousarray(entry_pos, dtype=np.float64)
    dirs = np.ascontiguousarray(entry_dir, dtype=np.float64)
    if pos.ndim != 2 or pos.shape[1] != 3 or dirs.shape != pos.shape:
        raise ValueError(f"entry_pos and entry_dir must both be [N,3]; got {pos.shape} and {dirs.shape}")
    n = pos.shape[0]
    if out is not None:
        exit_pos, exit_dir, status = out
        _check_out(exit_pos, (n, 3), np.float64, "out[0] (exit_pos)")
        _check_out(exit_dir, (n, 3), np.float64, "out[1] (exit_dir)")
        _check_out(status, (n,), np.int32, "out[2] (status)")
    else:
        exit_pos = np.empty((n, 3), dtype=np.float64)
        exit_dir = np.empty((n, 3), dtype=np.float64)
        status = np.empty(n, dtype=np.int32)
    counters = np.empty((2, n), dtype=np.int32) if return_counters else None
    p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    disk_xy = np.full((n, 2), np.nan, dtype=np.float64) if disk is not None else None
    extras = None
    if disk is not None or polyline is not None:
        extras = BhgExtras()
        if disk is not None:
            extras.disk_r_in, extras.disk_r_out, extras.disk_xy = float(disk[0]), float(disk[1]), disk_xy.ctypes.data
        if polyline is not None:
            if int(polyline) < 2 or lambda_max is None:
                raise ValueError("polyline needs K >= 2 samples and an explicit lambda_max")
            poly_xyz = np.full((n, int(polyline), 3), np.nan, dtype=np.float64)
            poly_count = np.zeros(n, dtype=np.int32)
            extras.poly_n, extras.poly_xyz, extras.poly_count = int(polyline), poly_xyz.ctypes.data, poly_count.ctypes.data
    _lib.check(lib.bhg_trace_schwarzschild_f64_host_ex(p(pos), p(dirs), p(exit_pos), p(exit_dir), p(status),
                                                       p(counters), n, ctypes.byref(params),
                                                       ctypes.byref(extras) if extras is not None else None,
                                                       int(device)))
    res = (exit_pos, exit_dir, status)
    if return_counters:
        res += (counters,)
    if disk is not None:
        res += (disk_xy,)
    if polyline is not None:
        res += (poly_xyz, poly_count)
    return res


def _trace_torch(entry_pos, entry_dir, params, return_counters, disk=None):
    import torch

    if not (entry_pos.is_cuda and entry_dir.is_cuda):
        raise ValueError("torch inputs must be CUDA tensors (pass numpy arrays for host data)")
    pos = entry_pos.to(torch.float64).contiguous()
    dirs = entry_dir.to(torch.float64).contiguous()
    if pos.ndim != 2 or pos.shape[1] != 3 or dirs.shape != pos.shape:
        raise ValueError("entry_pos and entry_dir must both be [N,3]")
    n = pos.shape[0]
    dev = pos.device
    exit_pos = torch.empty_like(pos)
    exit_dir = torch.empty_like(pos)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    counters = torch.empty((2, n), dtype=torch.int32, device=dev) if return_counters else None
    disk_xy = torch.full((n, 2), float("nan"), dtype=torch.float64, device=dev) if disk is not None else None
    extras = BhgExtras(float(disk[0]), float(disk[1]), disk_xy.data_ptr()) if disk is not None else None
    trace_device(pos.data_ptr(), dirs.data_ptr(), exit_pos.data_ptr(), exit_dir.data_ptr(), status.data_ptr(),
                 counters.data_ptr() if counters is not None else None, None, n, LAYOUT_AOS, params,
                 dev.index or 0, torch.cuda.current_stream(dev).cuda_stream, extras)
    res = (exit_pos, exit_dir, status)
    if return_counters:
        res += (counters,)
    if disk is not None:
        res += (disk_xy,)
    return res


def trace_device(in_ptr, in_dir_ptr, out_ptr, out_dir_ptr, status_ptr, counters_ptr, order_ptr, n, layout,
                 params: BhgParams, device=0, stream=0, extras: BhgExtras = None):
    """Raw device-pointer call (asynchronous on `stream`); the benchmark's hot call."""
    lib = _lib.load()
    if extras is None:
        _lib.check(lib.bhg_trace_schwarzschild_f64(in_ptr, in_dir_ptr, out_ptr, out_dir_ptr, status_ptr,
                                                   counters_ptr, order_ptr, int(n), int(layout),
                                                   ctypes.byref(params), int(device), stream or None))
    else:
        _lib.check(lib.bhg_trace_schwarzschild_f64_ex(in_ptr, in_dir_ptr, out_ptr, out_dir_ptr, status_ptr,
                                                      counters_ptr, order_ptr, int(n), int(layout),
                                                      ctypes.byref(params), ctypes.byref(extras), int(device),
                                                      stream or None))


def make_camera(origin, rotation, width, height, fov_x=0.6, fov_y=0.6, seed=42, jitter="philox", first_ray=0):
    """`struct bhg_camera` for the fused generate+trace entry points: the pinhole camera of the reference's
    render loop (RelativisticRenderEngine.py:181-189,223-230).  `origin` is relative to the black-hole centre,
    `rotation` the 3x3 camera-to-world matrix (raygen.look_at_rotation / euler_xyz_rotation)."""
    if jitter not in ("none", "philox"):
        raise ValueError("device-side generation supports jitter 'none' or 'philox' (the MT19937 stream of the "
                         "reference is sequential; use raygen.camera_rays for it)")
    cam = BhgCamera()
    cam.origin[:] = [float(v) for v in origin]
    cam.rotation[:] = [float(v) for v in np.asarray(rotation, dtype=np.float64).reshape(9)]
    cam.fov_x, cam.fov_y = float(fov_x), float(fov_y)
    cam.first_ray, cam.seed = int(first_ray), int(seed)
    cam.width, cam.height = int(width), int(height)
    cam.jitter, cam.reserved = (1 if jitter == "philox" else 0), 0
    return cam


def generate_rays(cam: BhgCamera, n, r_sphere, device=0):
    """Device-side primary rays: torch CUDA tensors (entry_pos[n,3], entry_dir[n,3], hit[n] int32: 0 or 5)."""
    import torch

    dev = torch.device("cuda", device)
    pos = torch.empty((n, 3), dtype=torch.float64, device=dev)
    d = torch.empty((n, 3), dtype=torch.float64, device=dev)
    hit = torch.empty(n, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().bhg_generate_rays_f64(ctypes.byref(cam), float(r_sphere), int(n), pos.data_ptr(),
                                                 d.data_ptr(), hit.data_ptr(), int(device),
                                                 torch.cuda.current_stream(dev).cuda_stream or None))
    return pos, d, hit


def trace_camera(cam: BhgCamera, n, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, *, max_step=math.inf,
                 eps_horizon=0.01, lambda_max=None, mode="parity", refill_threshold=0, device=0, want_pos=True,
                 return_counters=False, out="numpy", buffers=None):
    """Fused device-side ray generation + trace of `n` rays of `cam` (one frame or tile from ~150 bytes of input).

    out="numpy": host arrays (through bhg_trace_camera_f64_host; `buffers` may supply pre-allocated, e.g. pinned,
    (exit_pos|None, exit_dir, status) arrays).  out="torch": CUDA tensors on `device`, asynchronous.
    want_pos=False skips the exit positions (directions + status are what the RRE / CAM consumers read).
    Returns (exit_pos or None, exit_dir, status[, counters])."""
    params = make_params(M, r_sphere, rtol, atol, max_step, eps_horizon, lambda_max, mode, refill_threshold)
    lib = _lib.load()
    n = int(n)
    if out == "torch":
        import torch

        dev = torch.device("cuda", device)
        ep = torch.empty((n, 3), dtype=torch.float64, device=dev) if want_pos else None
        ed = torch.empty((n, 3), dtype=torch.float64, device=dev)
        st = torch.empty(n, dtype=torch.int32, device=dev)
        cnt = torch.empty((2, n), dtype=torch.int32, device=dev) if return_counters else None
        _lib.check(lib.bhg_trace_camera_f64(ctypes.byref(cam), ep.data_ptr() if want_pos else None, ed.data_ptr(),
                                            st.data_ptr(), cnt.data_ptr() if return_counters else None, n,
                                            ctypes.byref(params), int(device),
                                            torch.cuda.current_stream(dev).cuda_stream or None))
    else:
        if buffers is not None:
            ep, ed, st = buffers
            if want_pos:
                _check_out(ep, (n, 3), np.float64, "buffers[0] (exit_pos)")
            _check_out(ed, (n, 3), np.float64, "buffers[1] (exit_dir)")
            _check_out(st, (n,), np.int32, "buffers[2] (status)")
        else:
            ep = np.empty((n, 3), dtype=np.float64) if want_pos else None
            ed = np.empty((n, 3), dtype=np.float64)
            st = np.empty(n, dtype=np.int32)
        cnt = np.empty((2, n), dtype=np.int32) if return_counters else None
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(lib.bhg_trace_camera_f64_host(ctypes.byref(cam), p(ep) if want_pos else None, p(ed), p(st), p(cnt),
                                                 n, ctypes.byref(params), int(device)))
    if return_counters:
        return ep, ed, st, cnt
    return ep, ed, st


def trace_camera_f32(cam: BhgCamera, n, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, *, max_step=math.inf,
                     eps_horizon=0.01, lambda_max=None, refill_threshold=0, device=0, want_pos=False, buffers=None):
    """`trace_camera` with float32 host outputs (bhg_trace_camera_f32_host): FP64 integration, exit states rounded once
    to float32.  With want_pos=False (default) 16 bytes per ray come back - exit_dir + status, what the RRE / CAM
    consumers read (RelativisticRenderEngine.py:246, RelativisticRenderEngineCamEdition.py:228).
    `buffers` = (exit_pos f32 [n,3] | None, exit_dir f32 [n,3], status i32 [n]) C-contiguous (e.g. pinned_empty).
    Returns (exit_pos or None, exit_dir, status)."""
    params = make_params(M, r_sphere, rtol, atol, max_step, eps_horizon, lambda_max, "parity", refill_threshold)
    n = int(n)
    if buffers is not None:
        ep, ed, st = buffers
        _check_out(ed, (n, 3), np.float32, "exit_dir")
        _check_out(st, (n,), np.int32, "status")
        if want_pos:
            _check_out(ep, (n, 3), np.float32, "exit_pos")
    else:
        ep = np.empty((n, 3), np.float32) if want_pos else None
        ed, st = np.empty((n, 3), np.float32), np.empty(n, np.int32)
    p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(_lib.load().bhg_trace_camera_f32_host(ctypes.byref(cam), p(ep) if want_pos else None, p(ed), p(st), n,
                                                     ctypes.byref(params), int(device)))
    return (ep if want_pos else None), ed, st


def trace_f32(entry_pos, entry_dir, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, *, max_step=math.inf,
              eps_horizon=0.01, lambda_max=None, mode="parity", refill_threshold=0, image_width=0, device=0, out=None):
    """`trace` with float32 [N,3] host arrays in and out (Blender's native precision): FP64 integration of the exactly
    widened inputs, results rounded once to float32; 28 B/ray over PCIe instead of 100.  Returns
    (exit_pos f32, exit_dir f32, status i32)."""
    params = make_params(M, r_sphere, rtol, atol, max_step, eps_horizon, lambda_max, mode, refill_threshold,
                         image_width)
    pos = np.ascontiguousarray(entry_pos, dtype=np.float32)
    dirs = np.ascontiguousarray(entry_dir, dtype=np.float32)
    if pos.ndim != 2 or pos.shape[1] != 3 or dirs.shape != pos.shape:
        raise ValueError(f"entry_pos and entry_dir must both be [N,3]; got {pos.shape} and {dirs.shape}")
    n = pos.shape[0]
    if out is not None:
        exit_pos, exit_dir, status = out
        _check_out(exit_pos, (n, 3), np.float32, "out[0] (exit_pos)")
        _check_out(exit_dir, (n, 3), np.float32, "out[1] (exit_dir)")
        _check_out(status, (n,), np.int32, "out[2] (status)")
    else:
        exit_pos, exit_dir = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        status = np.empty(n, np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(_lib.load().bhg_trace_schwarzschild_f32io_host(p(pos), p(dirs), p(exit_pos), p(exit_dir), p(status), n,
                                                              ctypes.byref(params), int(device)))
    return exit_pos, exit_dir, status


def sky_uv(exit_dir, status=None):
    """Equirectangular sky-lookup coordinates of exit directions (the arithmetic of the reference's
    background_hit, RelativisticRenderEngine.py:366-378): float32 [N,2] = (-phi, 2 theta - 1); NaN where `status`
    says captured / start-inside / failed.  torch CUDA tensors in -> torch out."""
    import torch

    d = exit_dir.contiguous()
    n = d.shape[0]
    uv = torch.empty((n, 2), dtype=torch.float32, device=d.device)
    _lib.check(_lib.load().bhg_sky_uv_f32(d.data_ptr(), status.data_ptr() if status is not None else None, n,
                                          uv.data_ptr(), d.device.index or 0,
                                          torch.cuda.current_stream(d.device).cuda_stream or None))
    return uv


def trace_camera_sky(cam: BhgCamera, n, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, *, max_step=math.inf,
                     eps_horizon=0.01, lambda_max=None, mode="parity", refill_threshold=0, device=0, buffers=None):
    """Camera -> (uv[n,2] float32, status[n]) on the host: generate, trace and map on the device, copy back 12 B/ray."""
    params = make_params(M, r_sphere, rtol, atol, max_step, eps_horizon, lambda_max, mode, refill_threshold)
    n = int(n)
    uv, st = buffers if buffers is not None else (np.empty((n, 2), np.float32), np.empty(n, np.int32))
    if buffers is not None:
        _check_out(uv, (n, 2), np.float32, "buffers[0] (uv)")
        _check_out(st, (n,), np.int32, "buffers[1] (status)")
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(_lib.load().bhg_trace_camera_sky_host(ctypes.byref(cam), p(uv), p(st), n, ctypes.byref(params),
                                                     int(device)))
    return uv, st


def sum_counters(counters_ptr, status_ptr, n, device=0, stream=0):
    lib = _lib.load()
    a, b, c = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check(lib.bhg_sum_counters(counters_ptr, status_ptr, int(n), int(device), stream or None,
                                    ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
    return a.value, b.value, c.value


def device_numa_cpus(device=0):
    """CPUs of the NUMA node the GPU's PCIe root hangs off (sysfs), or None when the platform does not say."""
    lib = _lib.load()
    buf = ctypes.create_string_buffer(32)
    _lib.check(lib.bhg_device_pci_bus_id(int(device), buf, 32))
    try:
        with open(f"/sys/bus/pci/devices/{buf.value.decode().lower()}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
    except (OSError, ValueError):
        return None
    cpus = set()
    for part in spec.split(","):
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus or None


class _near_device:
    """Temporarily run the calling thread on the GPU's NUMA node so first-touch places new pages there."""

    def __init__(self, device):
        self.device, self.prev = device, None

    def __enter__(self):
        if self.device is None or not hasattr(os, "sched_setaffinity") or os.environ.get("BHG_NUMA_BIND") == "0":
            return self
        cpus = device_numa_cpus(self.device)
        allowed = os.sched_getaffinity(0)
        if cpus and (cpus & allowed):
            self.prev = allowed
            os.sched_setaffinity(0, cpus & allowed)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            os.sched_setaffinity(0, self.prev)
        return False


def pinned_empty(shape, dtype=np.float64, device=None):
    """numpy array backed by page-locked host memory (fast, asynchronous staging in the host entry point).
    With `device`, the pages are placed on the NUMA node of that GPU (matters when several ranks share a host)."""
    lib = _lib.load()
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(count * dtype.itemsize, 1)
    with _near_device(device):
        ptr = lib.bhg_host_alloc(nbytes)   # cudaHostAlloc populates and pins the pages here, on this node
    if not ptr:
        raise MemoryError("bhg_host_alloc failed: " + lib.bhg_last_error_string().decode())
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = (ptr, buf)
    return arr


_PINNED: dict = {}


def pinned_free(arr):
    key = arr.__array_interface__["data"][0]
    ptr, _ = _PINNED.pop(key)
    _lib.load().bhg_host_free(ptr)


def launch_count():
    return int(_lib.load().bhg_launch_count())


def selftest(device=0):
    out = (ctypes.c_double * 8)()
    rc = _lib.load().bhg_selftest(int(device), out)
    return rc, list(out)


def fp64_peak_tflops(device=0):
    tf, mhz = ctypes.c_double(0), ctypes.c_double(0)
    _lib.check(_lib.load().bhg_fp64_peak_tflops(int(device), ctypes.byref(tf), ctypes.byref(mhz)))
    return tf.value, mhz.value
