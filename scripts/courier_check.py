import sys; sys.path.insert(0, '.')
import numpy as np, torch
from blackhole_geodesic_calculator_b200 import api, raygen, distributed
pos, d = raygen.config_bundle(48, 48, 1, fov=0.55)
tp, td = torch.from_numpy(pos).cuda(), torch.from_numpy(d).cuda()
for n_, w_ in ((tp.shape[0], 48), (8192 + 77, 0), (tp.shape[0], 0)):
    p_, q_ = tp[:n_].contiguous(), td[:n_].contiguous()
    if n_ > tp.shape[0]:
        p_, q_ = tp.repeat(5, 1)[:n_].contiguous(), td.repeat(5, 1)[:n_].contiguous()
    fr = distributed.PeerFrame(n_)
    for t in fr.tensors()[:2]: t.fill_(float('nan'))
    fr.tensors()[2].fill_(-7)
    torch.cuda.synchronize()
    got = distributed.trace_sharded_peer(p_, q_, fr, route="courier", image_width=w_)
    torch.cuda.synchronize()
    ref = api.trace(p_, q_)
    bad = [int((a != b).sum()) for a, b in zip(got, ref)]
    nan = [int(torch.isnan(a).sum()) for a in got[:2]]
    print("courier", n_, w_, "mismatches", bad, "nan left", nan, "unwritten status", int((got[2] == -7).sum()))
    fr.close()
