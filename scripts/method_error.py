"""How accurate is the reference's method itself at its default tolerances?  C restatement at rtol=1e-3/atol=1e-6
against the same method converged (rtol=1e-12/atol=1e-15) on a 256x256 frame of the benchmark camera.  CPU only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import raygen  # noqa: E402
from oracle import port  # noqa: E402

pos, d = raygen.config_bundle(256, 256, 1, jitter="philox")
a = port.trace(pos, d)
b = port.trace(pos, d, rtol=1e-12, atol=1e-15)
esc = (a["status"] == 0) & (b["status"] == 0)
dd = np.linalg.norm(a["exit_dir"] - b["exit_dir"], axis=1)[esc]
print(f"{esc.sum()} escaped rays, {int((a['status'] != b['status']).sum())} status flips")
for q in (50, 90, 99, 99.9, 100):
    print(f"exit-direction error, percentile {q}: {np.percentile(dd, q):.2e} rad")
