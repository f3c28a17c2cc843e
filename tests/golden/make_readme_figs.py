"""Builds tests/golden/readme_fig5_fig6.npz from the ONLY known answers the reference holds for the geodesic path:
README Fig. 5 (images/large_impact_param_crossing.png, /root/reference/README.md:66-70: 17 rays from x = -15 R_s,
y = 3..19 R_s, initial direction +x) and Fig. 6 (images/small_impact_param.png, README.md:72-76: the same with rays
passing much closer).  Run in the build container (reads /root/reference; needs Pillow):

    python tests/golden/make_readme_figs.py

Stored: the pixel coordinates of the red trajectory lines of both figures, the axes mapping (matplotlib frame found
from the black axes box: columns 143..513, rows 58..427 <-> [-20, 20]^2 and [-5, 5]^2), the start definitions, and the
numbers read off the figures (exit ordinates on the right edge of Fig. 5, bottom-edge abscissae, ordinates of the ten
Fig. 6 rays on the left edge).  No reference source code is copied - only measurements of its published figures."""
import os

import numpy as np
from PIL import Image

REF = "/root/reference/images"
HERE = os.path.dirname(os.path.abspath(__file__))


def red_mask(name):
    im = np.array(Image.open(os.path.join(REF, name)).convert("RGB")).astype(int)
    r, g, b = im[..., 0], im[..., 1], im[..., 2]
    blk = (r < 40) & (g < 40) & (b < 40)
    rows = np.nonzero(blk.sum(axis=1) > 300)[0]
    cols = np.nonzero(blk.sum(axis=0) > 300)[0]
    assert len(rows) == 2 and len(cols) == 2, (rows, cols)     # the axes box
    return (r > 180) & (g < 100) & (b < 100), (int(cols[0]), int(cols[1]), int(rows[0]), int(rows[1])), im.shape[:2]


def runs(v):
    idx = np.nonzero(v)[0]
    out, s, p = [], None, None
    for i in idx:
        if s is None:
            s = p = i
        elif i != p + 1:
            out.append(0.5 * (s + p))
            s = i
        p = i
    if s is not None:
        out.append(0.5 * (s + p))
    return np.array(out)


def main():
    out = {}
    for tag, name, lim in (("fig5", "large_impact_param_crossing.png", 20.0), ("fig6", "small_impact_param.png", 5.0)):
        red, (c0, c1, r0, r1), shape = red_mask(name)
        ry, rx = np.nonzero(red)
        out[tag + "_red_px"] = np.stack([rx, ry], axis=1).astype(np.int16)
        out[tag + "_frame"] = np.array([c0, c1, r0, r1], dtype=np.int32)     # columns <-> x in [-lim, lim]; rows <-> y in [lim, -lim]
        out[tag + "_lim"] = np.array(lim)
        out[tag + "_shape"] = np.array(shape, dtype=np.int32)
        X = lambda c: -lim + (c - c0) / (c1 - c0) * 2 * lim
        Y = lambda r: lim - (r - r0) / (r1 - r0) * 2 * lim
        if tag == "fig5":
            out["fig5_right_edge_col"] = np.array(c1 - 2)
            out["fig5_right_edge_y"] = Y(runs(red[:, c1 - 2]))         # 15 visible (one hidden under the blue ray)
            out["fig5_bottom_row"] = np.array(r1 - 2)
            out["fig5_bottom_x"] = X(runs(red[r1 - 2, :]))
            out["fig5_y0"] = np.arange(3.0, 19.5, 1.0)
            out["fig5_x0"] = np.array(-15.0)
        else:
            out["fig6_left_col"] = np.array(c0 + 3)
            out["fig6_left_y"] = Y(runs(red[:, c0 + 3]))
            out["fig6_bottom_row"] = np.array(r1 - 2)
            out["fig6_bottom_x"] = X(runs(red[r1 - 2, :]))
            out["fig6_y0"] = np.arange(2.0, 2.95, 0.1)                 # identified from the left-edge ordinates
            out["fig6_x0"] = np.array(-15.0)
    np.savez_compressed(os.path.join(HERE, "readme_fig5_fig6.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v if v.size < 20 else "")


if __name__ == "__main__":
    main()
