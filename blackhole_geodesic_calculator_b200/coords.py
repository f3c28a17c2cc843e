"""Isotropic <-> Schwarzschild-coordinate maps of positions and coordinate directions (host side, numpy).

The device applies the same map at the kernel boundary when `coords="isotropic"` (include/bhgeo.h enum bhg_coords,
csrc/trace_kernel.cuh iso_to_schw / schw_to_iso); these functions exist for callers that hold results in one chart
and need the other, and for the tests that pin the map against README Fig. 5 / Fig. 6 of the reference
(/root/reference/README.md:64-76: rays "traced by the curvedpy python package", which "uses the Schwarzschild metric in
cartesian coordinates", README.md:174 - the isotropic ones, as the figures show).

Isotropic radius rho and Schwarzschild radius r:  r = rho (1 + r_s / 4 rho)^2,  rho = (r - r_s/2 + sqrt(r (r - r_s))) / 2.
Angular coordinates are shared.  A coordinate tangent with radial / tangential parts (d_r, d_t) in the isotropic chart
has parts ((1 - a^2) d_r, (1 + a)^2 d_t), a = r_s / 4 rho, in the Schwarzschild chart; directions are re-normalised
(a constant affine rescaling per ray, irrelevant to the path).
"""
from __future__ import annotations

import numpy as np


def schwarzschild_radius(rho, M=1.0):
    """r(rho) for an isotropic radius rho."""
    rho = np.asarray(rho, dtype=np.float64)
    return rho * (1.0 + 0.5 * M / rho) ** 2


def isotropic_radius(r, M=1.0):
    """rho(r) for a Schwarzschild radius r >= r_s = 2 M."""
    r = np.asarray(r, dtype=np.float64)
    return 0.5 * (r - M + np.sqrt(np.maximum(r * (r - 2.0 * M), 0.0)))


def isotropic_to_schwarzschild(pos, direction, M=1.0):
    """(pos[N,3], dir[N,3]) in the isotropic Cartesian chart -> the same point and unit tangent in the
    Schwarzschild-coordinate Cartesian chart."""
    pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
    d = np.asarray(direction, dtype=np.float64).reshape(-1, 3)
    rho2 = np.sum(pos * pos, axis=1)
    a = 0.5 * M / np.sqrt(rho2)
    c = (-2.0 * a / (1.0 + a)) * np.sum(d * pos, axis=1) / rho2
    v = d + c[:, None] * pos
    return pos * ((1.0 + a) ** 2)[:, None], v / np.linalg.norm(v, axis=1, keepdims=True)


def schwarzschild_to_isotropic(pos, direction, M=1.0):
    """Inverse of `isotropic_to_schwarzschild`."""
    pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
    d = np.asarray(direction, dtype=np.float64).reshape(-1, 3)
    r2 = np.sum(pos * pos, axis=1)
    r = np.sqrt(r2)
    rho = isotropic_radius(r, M)
    a = 0.5 * M / rho
    c = (2.0 * a / (1.0 - a)) * np.sum(d * pos, axis=1) / r2
    v = d + c[:, None] * pos
    return pos * (rho / r)[:, None], v / np.linalg.norm(v, axis=1, keepdims=True)


def points_to_isotropic(points, M=1.0):
    """Positions only (e.g. trajectory polylines [...,3]) Schwarzschild chart -> isotropic chart."""
    points = np.asarray(points, dtype=np.float64)
    r = np.linalg.norm(points, axis=-1, keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        return points * (isotropic_radius(r, M) / r)
