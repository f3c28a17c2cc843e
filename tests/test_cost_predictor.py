"""CPU test of the step-count predictor behind the queue-order features (csrc/trace_kernel.cuh estimate_attempts,
HOT_ESTIMATE): restated in numpy and held against the REAL scipy step counts stored in the golden sets
(attempts = (nfev - 2) / 6).  The predictor only orders the queue - it can never change a result - but the long-rays-
first list and the cost classes are only worth their passes if it keeps finding the long rays."""
import re
import os

import numpy as np

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def estimate_attempts(pos, d, rs):
    x, k = pos.astype(np.float32), d.astype(np.float32)
    L = np.cross(x, k)
    l2 = (L * L).sum(1)
    r2 = (x * x).sum(1)
    xk = (x * k).sum(1)
    r = np.sqrt(r2)
    b2 = l2 * r2 / (xk * xk + (1.0 - np.float32(rs) / r) * l2)
    sgn = b2 * np.float32(1.0 / (6.75 * rs * rs)) - 1.0
    u = np.abs(sgn)
    nz = np.abs(L[:, 2]) / np.sqrt(l2)
    lg = 1.0 - np.log2(np.maximum(u, 1e-7))
    base = np.where(u < 2.0, np.where(sgn > 0.0, 19.8 + 2.2 * lg, 33.3 + 1.2 * lg), 12.0)
    return base + np.where(nz < 0.5, (2.26 + 0.158 * (base - 12.0)) * (-1.0 - np.log2(np.maximum(nz, 1e-7))), 0.0)


def kernel_constants():
    src = open(os.path.join(ROOT, "blackhole_geodesic_calculator_b200", "csrc", "trace_kernel.cuh")).read()
    hot = float(re.search(r"constexpr float HOT_ESTIMATE = ([0-9.]+)f", src).group(1))
    for c in ("19.8f + 2.2f * lg", "33.3f + 1.2f * lg", "2.26f + 0.158f * (base - 12.0f)"):
        assert c in src, f"the numpy restatement no longer matches the kernel: {c}"
    return hot


def test_predictor_finds_the_long_rays_of_a_camera_frame():
    hot = kernel_constants()
    g = load_golden("cfg1_64x64.npz")
    integ = g["status"] != 2
    att = (g["nfev"][integ] - 2) / 6.0
    est = estimate_attempts(g["entry_pos"][integ], g["entry_dir"][integ], 2.0 * float(g["M"]))
    assert np.sqrt(((est - att) ** 2).mean()) < 6.0
    flagged = est >= hot
    assert flagged.mean() < 0.05                                     # a short list
    long_rays = att > 45
    assert long_rays.sum() >= 3 and flagged[long_rays].mean() >= 0.9    # that holds the tail of the launch
    assert att[~flagged].max() <= 60
    # cost classes of 4 attempts: rays of one class differ far less than rays of the frame
    cls = (est / 4.0).astype(int)
    within = sum((cls == c).sum() * att[cls == c].var() for c in np.unique(cls)) / len(att)
    assert 1.0 - within / att.var() > 0.5            # the classes explain more than half of the variance of the counts


def test_predictor_orders_a_near_critical_bundle():
    kernel_constants()
    g = load_golden("cfg5_nearcrit_3d.npz")
    integ = g["status"] != 2
    att = (g["nfev"][integ] - 2) / 6.0
    est = estimate_attempts(g["entry_pos"][integ], g["entry_dir"][integ], 2.0 * float(g["M"]))
    assert np.corrcoef(est, att)[0, 1] > 0.8
    # static lock-step efficiency of 32-ray groups: given order vs classes of the estimate
    def simt(order):
        a = att[order]
        m = len(a) // 32 * 32
        return a[:m].sum() / (32 * a[:m].reshape(-1, 32).max(1).sum())
    assert simt(np.argsort(-(est / 4.0).astype(int), kind="stable")) > simt(np.arange(len(att))) + 0.1
