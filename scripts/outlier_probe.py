"""Full frame, every 16th ray: rays where the CUDA path and the C restatement differ most; dumps them with the
scipy oracle's answer for the same rays."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import api, raygen  # noqa: E402
from oracle import port, schwarzschild_ref as ref  # noqa: E402

pos, d = raygen.config_bundle(1024, 1024, 5, jitter="philox")
sel = np.arange(0, pos.shape[0], int(sys.argv[1]) if len(sys.argv) > 1 else 16)
p, q = np.ascontiguousarray(pos[sel]), np.ascontiguousarray(d[sel])
ep, ed, st, cnt = api.trace(p, q, return_counters=True)
o = port.trace(p, q)
b = raygen.conserved_impact_parameter(p, q, 1.0)
dpos = np.linalg.norm(ep - o["exit_pos"], axis=1) / 60.0
ddir = np.linalg.norm(ed - o["exit_dir"], axis=1)
same = (st == o["status"])
dev = np.where(same & (st == 0), np.maximum(dpos, ddir), 0.0)
worst = np.argsort(-dev)[:12]
rows = []
for i in worst:
    s = ref.trace(p[i:i + 1], q[i:i + 1], 1.0, 60.0, 1e-3, 1e-6)
    theta_in = float(np.degrees(np.arccos(p[i, 2] / 60.0)))
    rows.append(dict(i=int(sel[i]), b=float(b[i]), dev=float(dev[i]), gpu_att=int(cnt[0][i]), port_att=int(o["n_attempt"][i]),
                     gpu_acc=int(cnt[1][i]), port_acc=int(o["n_accept"][i]), scipy_acc=int(s[4][0]), scipy_nfev=int(s[3][0]),
                     gpu_vs_scipy=float(np.linalg.norm(ep[i] - s[0][0]) / 60.0), port_vs_scipy=float(np.linalg.norm(o["exit_pos"][i] - s[0][0]) / 60.0),
                     theta_entry_deg=theta_in, exit_theta_deg=float(np.degrees(np.arccos(ep[i, 2] / 60.0)))))
print(json.dumps(dict(n=int(len(sel)), status_mismatch=int((~same).sum()), over_1e6=int((dev > 1e-6).sum()),
                      over_1e8=int((dev > 1e-8).sum()), attempts_differ=int((cnt[0] != o["n_attempt"]).sum()), worst=rows), indent=1))
