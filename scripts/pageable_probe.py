"""e2e through api.trace with ordinary (pageable) numpy arrays vs pinned ones."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from blackhole_geodesic_calculator_b200 import api, raygen
pos, d = raygen.config_bundle(1024, 1024, 5, jitter="philox")
n = pos.shape[0]
def wall(fn, reps=3):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
print("pageable numpy in/out (api.trace allocates outputs): %.1f ms" % wall(lambda: api.trace(pos, d, image_width=1024)))
ppos, pd = api.pinned_empty((n, 3)), api.pinned_empty((n, 3)); ppos[:] = pos; pd[:] = d
print("pinned inputs, pageable outputs: %.1f ms" % wall(lambda: api.trace(ppos, pd, image_width=1024)))
outs = (np.empty((n, 3)), np.empty((n, 3)), np.empty(n, np.int32))
print("pageable inputs, reused pageable outputs (out=): %.1f ms" % wall(lambda: api.trace(pos, d, image_width=1024, out=outs)))
pouts = (api.pinned_empty((n, 3)), api.pinned_empty((n, 3)), api.pinned_empty((n,), np.int32))
print("all pinned (out=): %.1f ms" % wall(lambda: api.trace(ppos, pd, image_width=1024, out=pouts)))
