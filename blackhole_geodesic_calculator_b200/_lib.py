"""ctypes loader for the C-ABI shared library (include/bhgeo.h).

Deliberately torch-free so it imports in Blender's bundled CPython (numpy only).  There is no CPU
fallback: if the library is missing or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BHG_LIB selects an alternative build of the same ABI (kernel tuning experiments); default is the in-tree build
LIB_PATH = os.environ.get("BHG_LIB") or os.path.join(_HERE, "lib", "libbhgeo.so")


class BhgParams(ctypes.Structure):
    """Mirror of `struct bhg_params` (include/bhgeo.h)."""
    _fields_ = [
        ("M", ctypes.c_double),
        ("r_sphere", ctypes.c_double),
        ("rtol", ctypes.c_double),
        ("atol", ctypes.c_double),
        ("max_step", ctypes.c_double),
        ("eps_horizon", ctypes.c_double),
        ("lambda_max", ctypes.c_double),
        ("mode", ctypes.c_int32),
        ("refill_threshold", ctypes.c_int32),
        ("image_width", ctypes.c_int32),
        ("coords", ctypes.c_int32),
    ]


class BhgCamera(ctypes.Structure):
    """Mirror of `struct bhg_camera` (include/bhgeo.h)."""
    _fields_ = [
        ("origin", ctypes.c_double * 3),
        ("rotation", ctypes.c_double * 9),
        ("fov_x", ctypes.c_double),
        ("fov_y", ctypes.c_double),
        ("first_ray", ctypes.c_int64),
        ("seed", ctypes.c_uint64),
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("jitter", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class BhgExtras(ctypes.Structure):
    """Mirror of `struct bhg_extras` (include/bhgeo.h)."""
    _fields_ = [("disk_r_in", ctypes.c_double), ("disk_r_out", ctypes.c_double), ("disk_xy", ctypes.c_void_p),
                ("poly_n", ctypes.c_int32), ("reserved", ctypes.c_int32), ("poly_xyz", ctypes.c_void_p),
                ("poly_count", ctypes.c_void_p)]


class BhgError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"bhgeo error {code}: {message}")
        self.code = code


# every symbol include/bhgeo.h declares, with its ctypes signature
_P = ctypes.c_void_p
_SIGNATURES = {
    "bhg_default_params": (None, [ctypes.POINTER(BhgParams)]),
    "bhg_trace_schwarzschild_f64": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, ctypes.c_int64, ctypes.c_int32,
                                                   ctypes.POINTER(BhgParams), ctypes.c_int32, _P]),
    "bhg_trace_schwarzschild_f64_host": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, ctypes.c_int64,
                                                        ctypes.POINTER(BhgParams), ctypes.c_int32]),
    "bhg_trace_schwarzschild_f64_ex": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, ctypes.c_int64, ctypes.c_int32,
                                                      ctypes.POINTER(BhgParams), ctypes.POINTER(BhgExtras),
                                                      ctypes.c_int32, _P]),
    "bhg_trace_schwarzschild_f64_host_ex": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, ctypes.c_int64,
                                                           ctypes.POINTER(BhgParams), ctypes.POINTER(BhgExtras),
                                                           ctypes.c_int32]),
    "bhg_trace_schwarzschild_f32io": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, ctypes.c_int64,
                                                     ctypes.POINTER(BhgParams), ctypes.c_int32, _P]),
    "bhg_trace_schwarzschild_f32io_host": (ctypes.c_int, [_P, _P, _P, _P, _P, ctypes.c_int64,
                                                          ctypes.POINTER(BhgParams), ctypes.c_int32]),
    "bhg_generate_rays_f64": (ctypes.c_int, [ctypes.POINTER(BhgCamera), ctypes.c_double, ctypes.c_int64, _P, _P, _P,
                                             ctypes.c_int32, _P]),
    "bhg_trace_camera_f64": (ctypes.c_int, [ctypes.POINTER(BhgCamera), _P, _P, _P, _P, ctypes.c_int64,
                                            ctypes.POINTER(BhgParams), ctypes.c_int32, _P]),
    "bhg_trace_camera_f64_host": (ctypes.c_int, [ctypes.POINTER(BhgCamera), _P, _P, _P, _P, ctypes.c_int64,
                                                 ctypes.POINTER(BhgParams), ctypes.c_int32]),
    "bhg_sky_uv_f32": (ctypes.c_int, [_P, _P, ctypes.c_int64, _P, ctypes.c_int32, _P]),
    "bhg_trace_camera_sky_host": (ctypes.c_int, [ctypes.POINTER(BhgCamera), _P, _P, ctypes.c_int64,
                                                 ctypes.POINTER(BhgParams), ctypes.c_int32]),
    "bhg_host_alloc": (_P, [ctypes.c_int64]),
    "bhg_host_free": (None, [_P]),
    "bhg_device_alloc": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(_P)]),
    "bhg_device_free": (ctypes.c_int, [_P, ctypes.c_int32]),
    "bhg_ipc_export": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_char_p]),
    "bhg_ipc_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int32, ctypes.POINTER(_P)]),
    "bhg_ipc_close": (ctypes.c_int, [_P, ctypes.c_int32]),
    "bhg_copy_rows": (ctypes.c_int, [_P, ctypes.c_int64, _P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_int32, _P]),
    "bhg_trace_camera_f32_host": (ctypes.c_int, [ctypes.POINTER(BhgCamera), _P, _P, _P, ctypes.c_int64,
                                                 ctypes.POINTER(BhgParams), ctypes.c_int32]),
    "bhg_trace_frame_shard_f64": (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.c_int64, _P, _P, _P, ctypes.c_int64,
                                                 ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(BhgParams), ctypes.c_int32,
                                                 _P]),
    "bhg_stream_write32": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, _P]),
    "bhg_stream_wait_geq32": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, _P]),
    "bhg_device_pci_bus_id": (ctypes.c_int, [ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32]),
    "bhg_sum_counters": (ctypes.c_int, [_P, _P, ctypes.c_int64, ctypes.c_int32, _P,
                                        ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                        ctypes.POINTER(ctypes.c_int64)]),
    "bhg_launch_count": (ctypes.c_int64, []),
    "bhg_selftest": (ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]),
    "bhg_fp64_peak_tflops": (ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(ctypes.c_double),
                                            ctypes.POINTER(ctypes.c_double)]),
    "bhg_last_error_string": (ctypes.c_char_p, []),
    "bhg_version": (ctypes.c_int, []),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Load lib/libbhgeo.so; raises (loudly) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI and the header drifted apart
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().bhg_last_error_string()
        raise BhgError(rc, msg.decode("utf-8", "replace") if msg else "")
