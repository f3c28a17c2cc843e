"""How much does the queue ORDER matter?  Same frame (configs[1]), order array = identity / rolled so that the
shadow rows are fetched last / shadow rows first / random permutation.  Single GPU, device-resident."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import api, raygen  # noqa: E402

W, H, S = 1280, 1024, 4
cam = api.make_camera(raygen.CFG_CAMERA_POS, raygen.look_at_rotation(raygen.CFG_CAMERA_POS), W, H,
                      raygen.CFG_FOV, raygen.CFG_FOV, seed=raygen.CFG_SEED, jitter="philox")
n = S * W * H
pos, d, _ = api.generate_rays(cam, n, raygen.CFG_R_SPHERE)
params = api.make_params(M=raygen.CFG_M, r_sphere=raygen.CFG_R_SPHERE, rtol=1e-3, atol=1e-6)
out_p, out_d = torch.empty_like(pos), torch.empty_like(d)
st = torch.empty(n, dtype=torch.int32, device="cuda")
cnt = torch.empty(2, n, dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream


def run(order):
    ts = []
    for it in range(8):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        api.trace_device(pos.data_ptr(), d.data_ptr(), out_p.data_ptr(), out_d.data_ptr(), st.data_ptr(),
                         cnt.data_ptr(), None if order is None else order.data_ptr(), n, api.LAYOUT_AOS, params,
                         stream=stream)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[3:]))


res = {"natural(no order array)": run(None)}
ident = torch.arange(n, dtype=torch.int32, device="cuda")
res["identity"] = run(ident)
plane = W * H
res["shadow_rows_last"] = run(((ident.long() + plane // 2) % n).int())
att = cnt[0].long()
res["mean_attempts"] = float(att.float().mean())
res["max_attempts"] = int(att.max())
res["rays_over_100_attempts"] = int((att > 100).sum())
res["longest_first(oracle LPT by true attempts)"] = run(torch.argsort(att, descending=True).int())
res["longest_last"] = run(torch.argsort(att).int())
g = torch.Generator(device="cuda").manual_seed(1)
res["random_permutation"] = run(torch.randperm(n, device="cuda", generator=g).int())
# cheap LPT: rays with > 40 attempts first (what a b-based predictor could find), the rest in natural order
hot = att > 40
res["hot_first_then_natural"] = run(torch.cat([ident[hot], ident[~hot]]))
res["hot_fraction"] = float(hot.float().mean())
print(json.dumps(res))
