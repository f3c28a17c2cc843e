"""Kernel time of ONE rank's shard of the configs[1] frame for world = 1, 2, 4, 8 (single GPU, no communication):
the floor of single-frame strong scaling, to separate the kernel's own small-n tail from delivery costs."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import api, distributed as D, raygen  # noqa: E402

cam = api.make_camera(raygen.CFG_CAMERA_POS, raygen.look_at_rotation(raygen.CFG_CAMERA_POS), 1280, 1024,
                      raygen.CFG_FOV, raygen.CFG_FOV, seed=raygen.CFG_SEED, jitter="philox")
n = 4 * 1280 * 1024
pos, d, _ = api.generate_rays(cam, n, raygen.CFG_R_SPHERE)
kw = dict(M=raygen.CFG_M, r_sphere=raygen.CFG_R_SPHERE, rtol=1e-3, atol=1e-6)
out = {}
for world in (1, 2, 4, 8):
    band, mine, m, ok = D.band_plan(n, 0, world, 1280)
    idx = torch.from_numpy((mine[:, None] * band + np.arange(band)[None, :]).reshape(-1)).cuda()
    p, q = pos.index_select(0, idx).contiguous(), d.index_select(0, idx).contiguous()
    for width in (0, 1280):
        ts = []
        for it in range(8):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            api.trace(p, q, image_width=width, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[f"world{world}/w{width}"] = {"rays": m, "ms": float(np.median(ts[3:])), "ideal_ms": None}
base = out["world1/w1280"]["ms"]
for k, v in out.items():
    v["ideal_ms"] = base * v["rays"] / n
print(json.dumps(out))
