"""Isolate the fused camera kernel from the host pipeline (tuning probe)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from blackhole_geodesic_calculator_b200 import api, raygen
W, H, SPP = 1024, 1024, 5
n = W * H * SPP
cpos = raygen.CFG_CAMERA_POS
cam = api.make_camera(cpos, raygen.look_at_rotation(cpos), W, H, 0.6, 0.6, seed=42, jitter="philox")
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("camera path (generate+trace), device, all: %.3f ms" % timeit(lambda: api.trace_camera(cam, n, out="torch")))
print("camera path (generate+trace), device, dir: %.3f ms" % timeit(lambda: api.trace_camera(cam, n, out="torch", want_pos=False)))
pos, d, hit = api.generate_rays(cam, n, 60.0)
print("generate_rays kernel                 : %.3f ms" % timeit(lambda: api.generate_rays(cam, n, 60.0)))
print("AOS kernel on the same rays (tiles)  : %.3f ms" % timeit(lambda: api.trace(pos, d, image_width=W)))
uvb = api.pinned_empty((n, 2), np.float32); stb = api.pinned_empty((n,), np.int32)
def wall(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for c in (262144, 1048576, 5242880):
    os.environ["BHG_CHUNK_RAYS"] = str(c)
    print("host uv path, chunk %8d        : %.3f ms" % (c, wall(lambda: api.trace_camera_sky(cam, n, buffers=(uvb, stb)))))
