#!/bin/bash
# compute-sanitizer on a small trace through every kernel variant (memcheck + racecheck + initcheck)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from blackhole_geodesic_calculator_b200 import api, raygen
pos, d = raygen.config_bundle(48, 48, 1, fov=0.55)
for kw in (dict(), dict(mode="plane"), dict(disk=(6.0, 20.0)), dict(refill_threshold=5), dict(image_width=48)):
    out = api.trace(pos, d, return_counters=True, **kw)
    print(kw, np.bincount(out[2], minlength=6))
rot = raygen.look_at_rotation(raygen.CFG_CAMERA_POS)
cam = api.make_camera(raygen.CFG_CAMERA_POS, rot, 48, 40, 1.0, 1.0, jitter="philox")
print(np.bincount(api.trace_camera(cam, 48 * 40)[2], minlength=6), api.trace_camera_sky(cam, 48 * 40)[0].shape)
# polyline output, float32 I/O, peer-frame routes (single process)
res = api.trace(pos[:256], d[:256], lambda_max=200.0, polyline=17)
print("poly", res[3].shape, int(res[4].min()), int(res[4].max()))
f = api.trace_f32(pos.astype(np.float32), d.astype(np.float32))
print("f32", np.bincount(f[2], minlength=6))
import torch
from blackhole_geodesic_calculator_b200 import distributed
tp, td = torch.from_numpy(pos).cuda(), torch.from_numpy(d).cuda()
frame = distributed.PeerFrame(tp.shape[0])
for route, chunks, w in (("stores", 1, 48), ("copy", 3, 48), ("copy", 2, 0)):
    got = distributed.trace_sharded_peer(tp, td, frame, route=route, chunks=chunks, image_width=w)
    torch.cuda.synchronize()
    print(route, chunks, w, torch.bincount(got[2].long(), minlength=6).tolist())
frame.close()
# round 2: courier route (band counters + courier kernel + in-place band reads), ragged frame included; unordered
# near-critical bundle (cost binning, scatter, adaptive budget) and a small image-ordered one (long-ray list);
# float32 camera outputs; isotropic boundary chart
for n_, w_ in ((tp.shape[0], 48), (8192 + 77, 0)):
    p_, q_ = tp[:n_].contiguous(), td[:n_].contiguous()
    if n_ > tp.shape[0]:
        p_, q_ = tp.repeat(5, 1)[:n_].contiguous(), td.repeat(5, 1)[:n_].contiguous()
    fr = distributed.PeerFrame(n_)
    got = distributed.trace_sharded_peer(p_, q_, fr, route="courier", image_width=w_)
    torch.cuda.synchronize()
    ref = api.trace(p_, q_)
    print("courier", n_, w_, all(torch.equal(a, b) for a, b in zip(got, ref)))
    fr.close()
p5, d5, _ = raygen.near_critical_bundle(1 << 13, in_plane=False)
o5 = api.trace(p5, d5, return_counters=True)
print("binned", np.bincount(o5[2], minlength=6), int(o5[3][0].max()))
print("f32cam", np.bincount(api.trace_camera_f32(cam, 48 * 40)[2], minlength=6))
print("iso", np.bincount(api.trace(pos, d, coords="isotropic")[2], minlength=6))
PY
for tool in memcheck racecheck initcheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san.py > gpurun_out/r2v_sanitizer_$tool.log 2>&1; echo "$tool rc=$?" | tee -a gpurun_out/r2v_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=" gpurun_out/r2v_sanitizer_$tool.log | tail -2
done
