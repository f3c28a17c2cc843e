"""AoS vs SoA ray buffers on the config-2 frame (north-star item: SoA layout with coalesced, vectorised double2 access).

With the pre-pass the trace kernel's own loads ARE double2: prepared records are planes of double2 in queue order
(coalesced 16-byte loads).  What remains layout-dependent is the pre-pass's read of the entry state (48 B/ray) and the
store of the exit state (52 B/ray).  Prints ms per frame for both layouts, with and without the image hint."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import api, raygen

W, H, SPP = 1024, 1024, 5
n = W * H * SPP
dev = torch.device("cuda", 0)
cam = api.make_camera(raygen.CFG_CAMERA_POS, raygen.look_at_rotation(raygen.CFG_CAMERA_POS), W, H * SPP, raygen.CFG_FOV,
                      raygen.CFG_FOV, seed=raygen.CFG_SEED, jitter="philox")
cam.height = H
pos, d, _ = api.generate_rays(cam, n, raygen.CFG_R_SPHERE, device=0)
soa_in = torch.cat([pos.t().contiguous(), d.t().contiguous()], 0).contiguous()   # 6 planes
soa_out = torch.empty_like(soa_in)
op, od = torch.empty_like(pos), torch.empty_like(d)
st = torch.empty(n, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream

def timeit(fn, it=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

out = {}
for width in (W, 0):
    prm = api.make_params(M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, image_width=width)
    out[f"aos_w{width}_ms"] = timeit(lambda: api.trace_device(pos.data_ptr(), d.data_ptr(), op.data_ptr(), od.data_ptr(),
                                                              st.data_ptr(), None, None, n, api.LAYOUT_AOS, prm, 0, stream))
    st_a = st.clone()
    out[f"soa_w{width}_ms"] = timeit(lambda: api.trace_device(soa_in.data_ptr(), None, soa_out.data_ptr(), None,
                                                              st.data_ptr(), None, None, n, api.LAYOUT_SOA, prm, 0, stream))
    same = torch.equal(st, st_a) and torch.equal(soa_out[:3].t().contiguous().view(torch.int64), op.view(torch.int64)) \
        and torch.equal(soa_out[3:].t().contiguous().view(torch.int64), od.view(torch.int64))
    out[f"soa_equals_aos_w{width}"] = bool(same)
print(json.dumps(out))
