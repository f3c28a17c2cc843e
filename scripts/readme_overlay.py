"""Overlay of the CUDA path's trajectories on README Fig. 5 / Fig. 6 of the reference (eyeball evidence next to
tests/test_readme_figures.py).  Runs on the GPU box: reads only tests/golden/readme_fig5_fig6.npz (the red pixels of
the figures), writes gpurun_out/r2_readme_fig{5,6}_overlay.png: figure lines in light red, GPU polylines in blue
(isotropic chart) and grey (the same start values read in the Schwarzschild-radius chart)."""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blackhole_geodesic_calculator_b200 import api  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "readme_fig5_fig6.npz"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for tag in ("fig5", "fig6"):
    h, w = g[tag + "_shape"]
    c0, c1, r0, r1 = g[tag + "_frame"]
    lim = float(g[tag + "_lim"])
    img = np.full((h, w, 3), 255, np.uint8)
    img[r0, c0:c1] = img[r1, c0:c1] = 0
    img[r0:r1, c0] = img[r0:r1, c1] = 0
    red = g[tag + "_red_px"]
    img[red[:, 1], red[:, 0]] = (255, 170, 170)
    y0 = g[tag + "_y0"]
    pos = np.stack([np.full_like(y0, float(g[tag + "_x0"])), y0, np.zeros_like(y0)], axis=1)
    d = np.tile([1.0, 0.0, 0.0], (len(y0), 1))
    for chart, colour in (("schwarzschild", (150, 150, 150)), ("isotropic", (0, 0, 200))):
        ep, ed, st, poly, cnt = api.trace(pos, d, 0.5, np.inf, 1e-9, 1e-12, lambda_max=90.0, polyline=20001, coords=chart)
        for i in range(len(y0)):
            P = poly[i, :cnt[i]]
            m = (np.abs(P[:, 0]) < lim) & (np.abs(P[:, 1]) < lim)
            cx = np.round(c0 + (P[m, 0] + lim) / (2 * lim) * (c1 - c0)).astype(int)
            cy = np.round(r0 + (lim - P[m, 1]) / (2 * lim) * (r1 - r0)).astype(int)
            img[cy, cx] = colour
    Image.fromarray(img).save(os.path.join(ROOT, "gpurun_out", f"r2_readme_{tag}_overlay.png"))
    print("wrote", tag)
