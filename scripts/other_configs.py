"""BASELINE configs 3 and 5 on the GPU (parity-test cases, not bench lines): timings + attempt statistics."""
import os, sys, json
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

out = {}
# config 3: 1920x1080 frame, camera at 200 M, rays that miss the sphere dropped
p3, d3 = raygen.random_impact_bundle(None)
n3 = p3.shape[0]
tp, td = torch.from_numpy(p3).cuda(), torch.from_numpy(d3).cuda()
ep, ed, st, cnt = api.trace(tp, td, return_counters=True)
att = cnt[0].double().mean().item()
for mode in ("parity", "plane"):
    ms = timeit(lambda: api.trace(tp, td, mode=mode))
    out[f"cfg3_{mode}"] = {"rays": n3, "ms": ms, "rays_per_s": n3 / ms * 1e3}
out["cfg3_parity"]["attempts_per_ray"] = att
out["cfg3_parity"]["captured_frac"] = (st == 1).double().mean().item()
# config 5: 2^20 near-critical rays, random plane orientation (incoherent order) and in-plane copy
for name, inplane in (("cfg5_3d", False), ("cfg5_inplane", True)):
    p5, d5, b5 = raygen.near_critical_bundle(1 << 20, in_plane=inplane)
    tp, td = torch.from_numpy(p5).cuda(), torch.from_numpy(d5).cuda()
    ep, ed, st, cnt = api.trace(tp, td, return_counters=True)
    a = cnt[0].cpu().numpy()
    rec = {"rays": 1 << 20, "attempts_mean": float(a.mean()), "attempts_p99": float(np.percentile(a, 99)), "attempts_max": int(a.max()),
           "captured_frac": float((st == 1).double().mean().item())}
    for T in (32, 24, 16, 12, 8):
        ms = timeit(lambda: api.trace(tp, td, refill_threshold=T))
        rec[f"ms_T{T}"] = ms
    rec["ms_plane_T32"] = timeit(lambda: api.trace(tp, td, mode="plane"))
    rec["ms_plane_T12"] = timeit(lambda: api.trace(tp, td, mode="plane", refill_threshold=12))
    best = min(v for k, v in rec.items() if k.startswith("ms_T"))
    rec["best_rays_per_s"] = (1 << 20) / best * 1e3
    out[name] = rec
# RRE call shape (RelativisticRenderEngine.py:293-294): camera inside the curved region, no sphere, fixed affine
# length 50, M = 0.5; 1024 x 1024 rays, only the end state is consumed (RRE.py:307-308)
rot = raygen.look_at_rotation((12.0, -8.0, 4.0))
dr = raygen.camera_rays(1024, 1024, 1, 1.0, 1.0, rot, 42, "philox")
pr = np.tile([12.0, -8.0, 4.0], (dr.shape[0], 1))
tp, td = torch.from_numpy(pr).cuda(), torch.from_numpy(dr).cuda()
kw = dict(M=0.5, r_sphere=float("inf"), lambda_max=50.0)
ep, ed, st, cnt = api.trace(tp, td, return_counters=True, **kw)
rec = {"rays": dr.shape[0], "attempts_per_ray": cnt[0].double().mean().item(),
       "status_counts": torch.bincount(st.long(), minlength=6).tolist()}
for width in (0, 1024):
    ms = timeit(lambda: api.trace(tp, td, image_width=width, **kw))
    rec[f"ms_w{width}"] = ms
rec["rays_per_s"] = dr.shape[0] / min(rec["ms_w0"], rec["ms_w1024"]) * 1e3
out["rre_call_shape"] = rec
print(json.dumps(out, indent=1))
