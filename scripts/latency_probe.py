"""Per-call latency of the drop-in route that keeps the reference's per-ray loop (one call per ray) and of small
batches: what a maintainer gets by only swapping the solver object, before restructuring the loop."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import adapters, api, raygen  # noqa: E402

pos, d = raygen.config_bundle(64, 64, 1)
out = {}


def per_call(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


for n in (1, 32, 1024, 4096):
    p, q = np.ascontiguousarray(pos[:n]), np.ascontiguousarray(d[:n])
    dt = per_call(lambda: api.trace(p, q), 300 if n <= 1024 else 100)
    out[f"trace_n{n}"] = {"us_per_call": dt * 1e6, "rays_per_s": n / dt}
pp, qq = api.pinned_empty((1, 3)), api.pinned_empty((1, 3))
pp[:], qq[:] = pos[:1], d[:1]
o = (api.pinned_empty((1, 3)), api.pinned_empty((1, 3)), api.pinned_empty((1,), np.int32))
dt = per_call(lambda: api.trace(pp, qq, out=o), 300)
out["trace_n1_pinned_reused_out"] = {"us_per_call": dt * 1e6}

gi = adapters.GeodesicIntegratorSchwarzschild(mass=0.5, time_like=False)
x0, k0 = np.array([12.0, -8.0, 4.0]), np.array([-0.8, 0.5, -0.2])
for npts in (2, 100, 10000):
    dt = per_call(lambda: gi.calc_trajectory(k0, x0, curve_end=50, nr_points_curve=npts), 200)
    out[f"calc_trajectory_{npts}pts"] = {"us_per_call": dt * 1e6, "rays_per_s": 1 / dt}
sw = adapters.SchwarzschildGeodesic()
loc = pos[0] / 60.0 * 30.0
dt = per_call(lambda: sw.ray_trace(d[0], loc, ratio_obj_to_blackhole=30.0, exit_tolerance=0.2), 200)
out["lim_ray_trace_per_ray"] = {"us_per_call": dt * 1e6, "rays_per_s": 1 / dt}
print(json.dumps(out, indent=1))
