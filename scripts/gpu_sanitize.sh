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
PY
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?" | tee -a gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=" gpurun_out/sanitizer_$tool.log | tail -2
done
