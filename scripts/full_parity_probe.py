"""Full-size parity of configs 3 and 5 against the C restatement (every ray): status flips inside / outside the
+-1e-2 M band around b_crit, step-count equality, exit-state deviation apart from pole-grazing planes."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import api, raygen  # noqa: E402
from oracle import port  # noqa: E402

B_CRIT = 3.0 * np.sqrt(3.0)
out = {}
sets = {"cfg3_1920x1080": raygen.random_impact_bundle(None)}
p5, d5, _ = raygen.near_critical_bundle(1 << 20, in_plane=False)
sets["cfg5_3d"] = (p5, d5)
p5, d5, _ = raygen.near_critical_bundle(1 << 20, in_plane=True)
sets["cfg5_inplane"] = (p5, d5)
for name, (p, d) in sets.items():
    p, d = np.ascontiguousarray(p), np.ascontiguousarray(d)
    ep, ed, st, cnt = api.trace(p, d, return_counters=True)
    o = port.trace(p, d)
    b = raygen.conserved_impact_parameter(p, d, 1.0)
    band = np.abs(b - B_CRIT) <= 1e-2
    flips = st != o["status"]
    nrm = np.cross(p, d)
    pole = np.abs(nrm[:, 2]) / np.linalg.norm(nrm, axis=1) < 1e-2
    esc = (st == 0) & (o["status"] == 0)
    dev = np.maximum(np.abs(ep - o["exit_pos"]).max(axis=1) / 60.0, np.abs(ed - o["exit_dir"]).max(axis=1))
    rec = dict(rays=int(len(st)), in_band=int(band.sum()), status_flips_in_band=int((flips & band).sum()),
               status_flips_outside_band=int((flips & ~band).sum()),
               attempts_equal_frac=float((cnt[0] == o["n_attempt"]).mean()),
               accepted_equal_frac=float((cnt[1] == o["n_accept"]).mean()),
               escaped=int(esc.sum()), pole_plane_rays=int((esc & pole).sum()),
               max_dev_escaped_non_pole=float(dev[esc & ~pole].max(initial=0.0)),
               over_1e6_non_pole=int((dev[esc & ~pole] > 1e-6).sum()),
               max_dev_escaped_pole=float(dev[esc & pole].max(initial=0.0)),
               over_1e6_pole=int((dev[esc & pole] > 1e-6).sum()),
               attempts_mean=float(cnt[0].mean()), attempts_max=int(cnt[0].max()))
    if rec["over_1e6_non_pole"]:
        w = np.argsort(-np.where(esc & ~pole, dev, 0))[:5]
        rec["worst_non_pole"] = [dict(i=int(i), b=float(b[i]), dev=float(dev[i]), att=int(cnt[0][i]),
                                      nz=float(abs(nrm[i, 2]) / np.linalg.norm(nrm[i]))) for i in w]
    out[name] = rec
print(json.dumps(out, indent=1))
