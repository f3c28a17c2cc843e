import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from blackhole_geodesic_calculator_b200 import api, raygen
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
# full 1920x1080 frame from 200 M incl. missing rays kept as NaN entries so the image structure survives
cam = np.array([0.6, -0.64, 0.48]) * 200.0
rot = raygen.look_at_rotation(cam)
c = api.make_camera(cam, rot, 1920, 1080, 0.6, 0.6, jitter="philox")
pos, d, hit = api.generate_rays(c, 1920 * 1080, 60.0)
print(os.environ.get("BHG_LIB", "default")[-6:], "cfg3 full image with tiles: %.4f ms" % timeit(lambda: api.trace(pos, d, image_width=1920)))
