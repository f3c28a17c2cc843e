"""Where does a 1/8 shard of the config-2 frame lose its time?  One GPU emulates rank 0 of `world` ranks: the shard's
rays integrated into local buffers, for the two dealing schemes (32-ray groups via `order`, 8-row bands compacted)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import api, distributed as D, raygen

W, H, SPP = 1024, 1024, 5
n = W * H * SPP
dev = torch.device("cuda", 0)
cam = api.make_camera(raygen.CFG_CAMERA_POS, raygen.look_at_rotation(raygen.CFG_CAMERA_POS), W, H * SPP, raygen.CFG_FOV,
                      raygen.CFG_FOV, seed=raygen.CFG_SEED, jitter="philox")
cam.height = H
pos, d, _ = api.generate_rays(cam, n, raygen.CFG_R_SPHERE, device=0)
op, od = torch.empty_like(pos), torch.empty_like(d)
st = torch.empty(n, dtype=torch.int32, device=dev)
cnt = torch.empty((2, n), dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream

def timeit(fn, it=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

res = {}
for world in (1, 2, 4, 8):
    for width in (0, W):
        order = torch.from_numpy(D.shard_order(n, 0, world, width)).to(dev)
        m = order.numel()
        prm = api.make_params(M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6)
        f = lambda: api.trace_device(pos.data_ptr(), d.data_ptr(), op.data_ptr(), od.data_ptr(), st.data_ptr(), None,
                                     order.data_ptr(), m, api.LAYOUT_AOS, prm, 0, stream)
        res[f"groups/world{world}/w{width}"] = timeit(f)
    # bands of 8 rows dealt cyclically, compacted (copy route): contiguous local arrays with the tile hint
    band = 8 * W
    nb = n // band
    mine = np.arange(0, nb, world)
    idx = (mine[:, None] * band + np.arange(band)[None, :]).reshape(-1)
    ti = torch.from_numpy(idx).to(dev)
    p2, d2 = pos.index_select(0, ti).contiguous(), d.index_select(0, ti).contiguous()
    m = p2.shape[0]
    for width in (0, W):
        prm = api.make_params(M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, image_width=width)
        f = lambda: api.trace_device(p2.data_ptr(), d2.data_ptr(), op.data_ptr(), od.data_ptr(), st.data_ptr(), None,
                                     None, m, api.LAYOUT_AOS, prm, 0, stream)
        res[f"bands/world{world}/w{width}"] = timeit(f)
    # attempts statistics of the shard
    prm = api.make_params(M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, image_width=W)
    api.trace_device(p2.data_ptr(), d2.data_ptr(), op.data_ptr(), od.data_ptr(), st.data_ptr(), cnt.data_ptr(), None, m,
                     api.LAYOUT_AOS, prm, 0, stream)
    torch.cuda.synchronize()
    a = cnt.view(-1)[:m]
    res[f"attempts/world{world}"] = {"mean": float(a.double().mean()), "max": int(a.max()), "over100": int((a > 100).sum())}
print(json.dumps(res, indent=1))
