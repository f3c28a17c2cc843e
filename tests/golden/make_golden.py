"""Generates tests/golden/*.npz with the REAL scipy path (oracle/schwarzschild_ref.py: sympy-derived RHS +
scipy.integrate.solve_ivp RK45), i.e. the restated reference method — run in the build container:

    python tests/golden/make_golden.py

The reference itself (`/root/reference`) cannot be imported or run (bpy, mathutils, curvedpy absent), so
these vectors pin our CUDA path and the C port to the published method, not to the reference's bytes:
PARITY UNPINNED (see oracle/schwarzschild_ref.py).  scipy 1.18.1 / numpy 2.3.5 / sympy 1.14.0.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from blackhole_geodesic_calculator_b200 import raygen  # noqa: E402
from oracle import schwarzschild_ref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run(pos, d, procs=8, **kw):
    res = R.trace_pool(pos, d, procs, chunk=64, **kw)
    ep, ed, st, nfev, nacc = res[:5]
    out = dict(entry_pos=pos, entry_dir=d, exit_pos=ep, exit_dir=ed, status=st, nfev=nfev, n_accept=nacc)
    extra = list(res[5:])
    if "disk" in kw and kw["disk"] is not None:
        out["disk_xy"] = extra.pop(0)
    if "polyline" in kw and kw["polyline"] is not None:
        out["poly_xyz"], out["poly_count"] = extra.pop(0), extra.pop(0)
    return out


def save(name, data, **meta):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **data, **{k: np.asarray(v) for k, v in meta.items()})
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in data.items()},
          "status hist", np.bincount(data["status"], minlength=5))


def analytic_dphi(b, M=1.0, R_=60.0):
    """2 * int_{1/R}^{u_p} du / sqrt(1/b^2 - u^2 (1 - 2 M u))   (SURVEY.md A.5), mpmath 30 digits."""
    import mpmath as mp
    mp.mp.dps = 30
    b = mp.mpf(b)
    # 1/b^2 - u^2 + 2 M u^3 = 2 M (u - u_neg)(u - u_p)(u - u_3); with u = u_p - s^2 the integrand becomes
    # 2 / sqrt(2 M (u - u_neg)(u_3 - u)), free of the end-point singularity at the periapsis u_p
    roots = sorted(mp.re(r) for r in mp.polyroots([2 * M, -1, 0, 1 / b**2], maxsteps=500, extraprec=400))
    u_neg, up, u3 = roots
    g = lambda s: 2 / mp.sqrt(2 * M * ((up - s * s) - u_neg) * (u3 - (up - s * s)))
    return float(2 * mp.quad(g, [0, mp.sqrt(up - mp.mpf(1) / R_)]))


def main():
    # config 1: 64x64 pinhole bundle, exact reference generator (MT19937 seed 42), default tolerances
    pos, d = raygen.config_bundle(64, 64, 1, jitter="mt19937")
    save("cfg1_64x64.npz", run(pos, d), M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6)

    # config 5: near-critical sweep b in [5.0, 5.4] M, random plane orientation + in-plane copy
    p5, d5, b5 = raygen.near_critical_bundle(384, in_plane=False)
    g = run(p5, d5)
    g["b"] = b5
    save("cfg5_nearcrit_3d.npz", g, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6)
    p5, d5, b5 = raygen.near_critical_bundle(256, in_plane=True, seed=7)
    g = run(p5, d5)
    g["b"] = b5
    save("cfg5_nearcrit_plane.npz", g, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6)

    # config 3 sample: camera at 200 M, impact parameters 0..60 M (every 2000th ray of the 1920x1080 frame)
    p3, d3 = raygen.random_impact_bundle(None)
    sel = np.arange(0, p3.shape[0], 2000)
    save("cfg3_sample.npz", run(p3[sel], d3[sel]), M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6)

    # tight tolerances on a 16x16 bundle
    pt, dt = raygen.config_bundle(16, 16, 1, jitter="mt19937", fov=0.5)
    save("tight_16x16.npz", run(pt, dt, rtol=1e-9, atol=1e-12), M=1.0, r_sphere=60.0, rtol=1e-9, atol=1e-12)

    # RRE call shape: no sphere, fixed affine length 50, camera inside the curved region, M = 0.5
    # (RelativisticRenderEngine.py:293-294,506-508)
    rot = raygen.look_at_rotation((12.0, -8.0, 4.0))
    dr = raygen.camera_rays(16, 16, 1, 1.0, 1.0, rot, 42, "mt19937")
    pr = np.tile([12.0, -8.0, 4.0], (dr.shape[0], 1))
    save("rre_shape_16x16.npz", run(pr, dr, M=0.5, r_sphere=np.inf, lambda_max=50.0), M=0.5, r_sphere=np.inf,
         rtol=1e-3, atol=1e-6, lambda_max=50.0)

    # max_step-limited rays (author's cached cameras used small max_step, CamEdition.py:216)
    pm, dm = raygen.config_bundle(8, 8, 1, jitter="mt19937", fov=0.45)
    save("maxstep_8x8.npz", run(pm, dm, max_step=1.0), M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, max_step=1.0)

    # edge cases
    Rr = 60.0
    e_pos = [
        [1.5, 0.0, 0.0],                 # inside the horizon -> START_INSIDE_HOLE
        [2.005, 0.0, 0.0],               # between r_s and r_s + eps -> START_INSIDE_HOLE
        [-Rr, 0.0, 0.0],                 # radial infall, b = 0 (equatorial)
        [-Rr / math.sqrt(3)] * 3,        # radial infall, generic orientation
        [0.0, Rr, 0.0],                  # tangential entry: leaves immediately
        [0.0, 0.0, Rr],                  # on the +z axis (coordinate pole): rho = 0 -> non-finite -> STEP_FAILED
        [-math.sqrt(Rr**2 - 30.0**2), 30.0, 0.0],   # lambda exhausted (lambda_max below)
        [-math.sqrt(Rr**2 - 1.0), 0.6, 0.8],        # passes close to the pole axis region, captured
        [30.0, 10.0, -5.0],              # starts inside the sphere (camera inside), heads out
        [30.0, 10.0, -5.0],              # starts inside the sphere, heads in
    ]
    e_dir = [
        [1, 0, 0], [1, 0, 0], [1, 0, 0], [1 / math.sqrt(3)] * 3, [1, 0, 0], [0.6, 0.0, -0.8],
        [1, 0, 0], [1, 0, 0], [0.6, 0.64, 0.48], [-0.6, -0.64, -0.48],
    ]
    e_pos, e_dir = np.array(e_pos, float), np.array(e_dir, float)
    ge = run(e_pos, e_dir, procs=1, lambda_max=80.0)
    save("edge_cases.npz", ge, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, lambda_max=80.0)

    # equatorial disk crossing (checkHitDisk, LIM.py:413-438) as a non-terminal event: camera above the plane,
    # annulus 6 M .. 20 M, plus near-critical 3-D rays (multiple plane crossings before capture / escape)
    pd_, dd_ = raygen.config_bundle(40, 40, 1, jitter="mt19937", fov=0.55)
    p5, d5, _ = raygen.near_critical_bundle(192, in_plane=False, seed=11)
    gd = run(np.concatenate([pd_, p5]), np.concatenate([dd_, d5]), disk=(6.0, 20.0))
    save("disk_crossing.npz", gd, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, disk_r_in=6.0, disk_r_out=20.0)

    # polyline samples on linspace(0, curve_end, K) (t_eval of solve_ivp; RelativisticRenderEngine.py:293-294):
    # RRE call shape (no sphere, M = 0.5, curve_end = 50) and sphere contract (lambda_max = 200)
    rot = raygen.look_at_rotation((12.0, -8.0, 4.0))
    dr = raygen.camera_rays(8, 8, 1, 1.0, 1.0, rot, 42, "mt19937")
    pr = np.tile([12.0, -8.0, 4.0], (dr.shape[0], 1))
    save("polyline_rre.npz", run(pr, dr, M=0.5, r_sphere=np.inf, lambda_max=50.0, polyline=33), M=0.5, r_sphere=np.inf,
         rtol=1e-3, atol=1e-6, lambda_max=50.0, polyline=33)
    pp, dp = raygen.config_bundle(8, 8, 1, jitter="mt19937", fov=0.45)
    save("polyline_sphere.npz", run(pp, dp, lambda_max=200.0, polyline=41), M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6,
         lambda_max=200.0, polyline=41)

    # analytic known answers: deflection between sphere entry and exit
    bflat = np.array([5.3, 6.0, 8.0, 12.0, 20.0, 40.0])
    pk = np.stack([-np.sqrt(Rr**2 - bflat**2), bflat, np.zeros_like(bflat)], axis=1)
    dk = np.tile([1.0, 0.0, 0.0], (len(bflat), 1))
    bcons = raygen.conserved_impact_parameter(pk, dk, 1.0)
    dphi = np.array([analytic_dphi(b) for b in bcons])
    gk = run(pk, dk, procs=1, rtol=1e-12, atol=1e-14)
    gk["b_flat"] = bflat
    gk["b"] = bcons
    gk["dphi_analytic"] = dphi
    save("analytic_kat.npz", gk, M=1.0, r_sphere=60.0, rtol=1e-12, atol=1e-14)
    # report
    e_in = np.arctan2(pk[:, 1], pk[:, 0])
    e_out = np.arctan2(gk["exit_pos"][:, 1], gk["exit_pos"][:, 0])
    swept = np.mod(e_in - e_out, 2 * np.pi)
    for b, a, s in zip(bflat, dphi, swept):
        k = round((a - s) / (2 * np.pi))
        print(f"  b_flat {b:5.2f}  analytic {a:.9f}  rk45(1e-12) {s + 2 * np.pi * k:.9f}")


if __name__ == "__main__":
    main()
