"""Cost binning A/B (run once per BHG_BIN setting: the library reads it once): full config-2 frame with and without the
image hint, a 1/8 band shard, config 3, config 5 (random planes / in-plane); results must be bit-identical."""
import hashlib, json, os, sys
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
    return float(np.median(ts))

def digest(res):
    h = hashlib.sha256()
    for t in res[:3]:
        h.update(t.cpu().numpy().tobytes())
    return h.hexdigest()[:16]

out = {"BHG_BIN": os.environ.get("BHG_BIN", "1")}
W, H, SPP = 1024, 1024, 5
n = W * H * SPP
cam = api.make_camera(raygen.CFG_CAMERA_POS, raygen.look_at_rotation(raygen.CFG_CAMERA_POS), W, H * SPP, raygen.CFG_FOV,
                      raygen.CFG_FOV, seed=raygen.CFG_SEED, jitter="philox")
cam.height = H
pos, d, _ = api.generate_rays(cam, n, raygen.CFG_R_SPHERE, device=0)
for width in (W, 0):
    r = api.trace(pos, d, image_width=width, return_counters=True)
    out[f"cfg2_w{width}"] = {"ms": timeit(lambda: api.trace(pos, d, image_width=width)), "sha": digest(r),
                             "attempts": float(r[3][0].double().mean())}
band = 8 * W
idx = (np.arange(0, n // band, 8)[:, None] * band + np.arange(band)[None, :]).reshape(-1)
ti = torch.from_numpy(idx).cuda()
p2, d2 = pos.index_select(0, ti).contiguous(), d.index_select(0, ti).contiguous()
for width in (W, 0):
    r = api.trace(p2, d2, image_width=width)
    out[f"shard8_w{width}"] = {"ms": timeit(lambda: api.trace(p2, d2, image_width=width)), "sha": digest(r), "rays": p2.shape[0]}
p3, d3 = raygen.random_impact_bundle(None)
tp, td = torch.from_numpy(p3).cuda(), torch.from_numpy(d3).cuda()
r = api.trace(tp, td, return_counters=True)
out["cfg3"] = {"rays": p3.shape[0], "ms": timeit(lambda: api.trace(tp, td)), "sha": digest(r), "attempts": float(r[3][0].double().mean())}
for name, inplane in (("cfg5_3d", False), ("cfg5_inplane", True)):
    p5, d5, b5 = raygen.near_critical_bundle(1 << 20, in_plane=inplane)
    tp, td = torch.from_numpy(p5).cuda(), torch.from_numpy(d5).cuda()
    r = api.trace(tp, td, return_counters=True)
    out[name] = {"rays": 1 << 20, "ms": timeit(lambda: api.trace(tp, td)), "sha": digest(r), "attempts": float(r[3][0].double().mean())}
print(json.dumps(out))
