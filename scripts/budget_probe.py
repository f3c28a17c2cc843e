"""Idle-budget sweep of the refill policy on configs 2, 3, 5 (one process per setting: the library reads the env once)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import api, raygen

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return round(float(np.median(ts)), 4)

out = {"budget": os.environ.get("BHG_IDLE_BUDGET", "96")}
p2, d2 = raygen.config_bundle(1024, 1024, 2, jitter="philox")
tp, td = torch.from_numpy(p2).cuda(), torch.from_numpy(d2).cuda()
out["cfg2_2spp_w1024"] = timeit(lambda: api.trace(tp, td, image_width=1024))
p3, d3 = raygen.random_impact_bundle(None)
tp, td = torch.from_numpy(p3).cuda(), torch.from_numpy(d3).cuda()
out["cfg3"] = timeit(lambda: api.trace(tp, td))
for name, inplane in (("cfg5_3d", False), ("cfg5_inplane", True)):
    p5, d5, _ = raygen.near_critical_bundle(1 << 20, in_plane=inplane)
    tp, td = torch.from_numpy(p5).cuda(), torch.from_numpy(d5).cuda()
    out[name] = timeit(lambda: api.trace(tp, td))
print(json.dumps(out))
