"""CPU tests: the C-ABI library loads, exports every symbol include/bhgeo.h declares, validates arguments
without a GPU, and fails loudly (no CPU fallback) when no device is present."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

from blackhole_geodesic_calculator_b200 import _lib, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "bhgeo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bhg_[a-z0-9_]+)\s*\(", src)))


def test_header_and_loader_agree():
    assert header_functions() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), name


def test_version_and_defaults():
    lib = _lib.load()
    assert lib.bhg_version() == 120
    p = _lib.BhgParams()
    lib.bhg_default_params(ctypes.byref(p))
    assert (p.M, p.r_sphere, p.rtol, p.atol, p.eps_horizon, p.mode) == (1.0, 60.0, 1e-3, 1e-6, 0.01, 0)
    assert math.isinf(p.max_step) and p.lambda_max == 0.0


def test_params_struct_layout():
    assert ctypes.sizeof(_lib.BhgParams) == 7 * 8 + 4 * 4


@pytest.mark.parametrize("kw,frag", [
    (dict(M=-1.0), "M must be"),
    (dict(r_sphere=1.0), "r_sphere"),
    (dict(rtol=0.0), "rtol"),
    (dict(max_step=0.0), "max_step"),
    (dict(refill_threshold=33), "refill_threshold"),
    (dict(r_sphere=math.inf), "lambda_max must be given"),
])
def test_argument_validation_needs_no_gpu(kw, frag):
    pos = np.zeros((4, 3))
    with pytest.raises(_lib.BhgError) as e:
        api.trace(pos, pos, **kw)
    assert e.value.code == -1 and frag in str(e.value)


def test_shape_validation():
    with pytest.raises(ValueError):
        api.trace(np.zeros((4, 2)), np.zeros((4, 2)))
    with pytest.raises(ValueError):
        api.trace(np.zeros((4, 3)), np.zeros((5, 3)))
    with pytest.raises(ValueError):
        api.make_params(mode="fast")


def test_no_cpu_fallback():
    """Without a CUDA device the product path must raise, not compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    pos = np.array([[-60.0, 0.0, 0.0]])
    with pytest.raises(_lib.BhgError) as e:
        api.trace(pos, np.array([[1.0, 0.0, 0.0]]))
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "blackhole_geodesic_calculator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "rk45_port" not in text or f.endswith(".md"), f


def test_maximum_size_is_rejected_before_any_buffer_is_touched():
    """n beyond 2^31 - 1 rays per call is refused by validation (no device, no buffers needed)."""
    lib = _lib.load()
    p = api.make_params()
    rc = lib.bhg_trace_schwarzschild_f64(None, None, None, None, None, None, None, 2**31, 1, ctypes.byref(p), 0, None)
    assert rc == -1 and b"exceeds" in lib.bhg_last_error_string()
    rc = lib.bhg_trace_schwarzschild_f64_host(None, None, None, None, None, None, -5, ctypes.byref(p), 0)
    assert rc == -1 and b"negative" in lib.bhg_last_error_string()


def test_camera_validation_needs_no_gpu():
    import numpy as np
    cam = api.make_camera((120.0, -80.0, 40.0), np.eye(3), 0, 16)
    lib = _lib.load()
    p = api.make_params()
    rc = lib.bhg_trace_camera_f64_host(ctypes.byref(cam), None, None, None, None, 16, ctypes.byref(p), 0)
    assert rc == -1 and b"width/height" in lib.bhg_last_error_string()
    with pytest.raises(ValueError):
        api.make_camera((1, 2, 3), np.eye(3), 8, 8, jitter="mt19937")


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/bhgeo.h compiles as C99 and a C program can link the library and call it (the boundary is a C ABI,
    not a Python extension): defaults, struct sizes, and an argument error reported through the C error string."""
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if not gcc:
        pytest.skip("no C compiler")
    src = tmp_path / "use_bhgeo.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "bhgeo.h"
int main(void) {
    bhg_params p;
    bhg_default_params(&p);
    if (sizeof(bhg_params) != 72 || p.M != 1.0 || p.r_sphere != 60.0 || p.rtol != 1e-3 || p.atol != 1e-6) return 1;
    p.rtol = -1.0;
    double buf[3] = {0, 0, 0};
    int status = 0;
    int rc = bhg_trace_schwarzschild_f64_host(buf, buf, buf, buf, &status, NULL, 1, &p, 0);
    if (rc != BHG_ERR_INVALID_ARGUMENT || !strstr(bhg_last_error_string(), "rtol")) return 2;
    printf("version %d\n", bhg_version());
    return 0;
}
''')
    exe = tmp_path / "use_bhgeo"
    libdir = os.path.dirname(os.path.abspath(_lib.LIB_PATH))
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lbhgeo", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.startswith("version ")


def test_caller_buffers_are_checked_before_a_pointer_reaches_the_c_abi():
    """every entry point that takes caller arrays validates shape, dtype, contiguity (no GPU needed: the
    check precedes the library call)."""
    cam = api.make_camera((120.0, -80.0, 40.0), np.eye(3), 8, 8)
    n = 64
    good32, goodst = np.empty((n, 3), np.float32), np.empty(n, np.int32)
    with pytest.raises(ValueError, match="float32"):
        api.trace_camera_f32(cam, n, buffers=(None, np.empty((n, 3)), goodst))
    with pytest.raises(ValueError, match="status"):
        api.trace_camera(cam, n, buffers=(np.empty((n, 3)), np.empty((n, 3)), np.empty(n - 1, np.int32)))
    with pytest.raises(ValueError, match="uv"):
        api.trace_camera_sky(cam, n, buffers=(np.empty((n, 3), np.float32), goodst))
    with pytest.raises(ValueError, match="C-contiguous"):
        api.trace_f32(np.zeros((n, 3), np.float32), np.ones((n, 3), np.float32),
                      out=(np.empty((n, 6), np.float32)[:, ::2], good32, goodst))
    with pytest.raises(ValueError, match="exit_pos"):
        api.trace(np.zeros((n, 3)), np.ones((n, 3)), out=(np.empty((n, 3), np.float32), np.empty((n, 3)), goodst))
    with pytest.raises(ValueError, match="r_in"):
        api.trace(np.zeros((n, 3)), np.ones((n, 3)), disk=(20.0, 6.0))
    with pytest.raises(ValueError, match="r_in"):
        api.trace(np.zeros((n, 3)), np.ones((n, 3)), disk=(0.0, 0.0))
