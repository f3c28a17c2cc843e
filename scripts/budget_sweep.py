"""Tune the adaptive refill budget on a coherent (config 2) and an incoherent (config 5, 3-D) workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from blackhole_geodesic_calculator_b200 import api, raygen
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
p2, d2 = raygen.config_bundle(1024, 1024, 5, jitter="philox")
a2, b2 = torch.from_numpy(p2).cuda(), torch.from_numpy(d2).cuda()
p5, d5, _ = raygen.near_critical_bundle(1 << 20, in_plane=False)
a5, b5 = torch.from_numpy(p5).cuda(), torch.from_numpy(d5).cuda()
p3, d3 = raygen.random_impact_bundle(None)
a3, b3 = torch.from_numpy(p3).cuda(), torch.from_numpy(d3).cuda()
print("explicit T: cfg2 tiles T32 %.3f | cfg2 rows T32 %.3f | cfg5 T32 %.3f T8 %.3f | cfg3 T32 %.3f" % (
    timeit(lambda: api.trace(a2, b2, image_width=1024, refill_threshold=32)), timeit(lambda: api.trace(a2, b2, refill_threshold=32)),
    timeit(lambda: api.trace(a5, b5, refill_threshold=32)), timeit(lambda: api.trace(a5, b5, refill_threshold=8)),
    timeit(lambda: api.trace(a3, b3, refill_threshold=32))))
for B in (16, 24, 32, 48, 64, 96, 128, 192, 256):
    os.environ["BHG_IDLE_BUDGET"] = str(B)
    print("budget %3d: cfg2 tiles %.3f | cfg2 rows %.3f | cfg5 %.3f | cfg3 %.3f | cfg5 plane %.3f" % (B,
          timeit(lambda: api.trace(a2, b2, image_width=1024)), timeit(lambda: api.trace(a2, b2)),
          timeit(lambda: api.trace(a5, b5)), timeit(lambda: api.trace(a3, b3)), timeit(lambda: api.trace(a5, b5, mode="plane"))))
