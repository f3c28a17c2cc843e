"""Host-side primary-ray generation and sphere entry — the caller side of the hot path.

Restates (does not copy) the reference's pinhole generator
  /root/reference/raytracer/RelativisticRenderEngine.py:185-189  (aspect, dx, dy, random.seed)
  /root/reference/raytracer/RelativisticRenderEngine.py:195-230  (loop order s -> y -> x, x-jitter drawn
      before y-jitter, direction = R_cam . (x_render + jx, y_render + jy, -1), normalised)
and the sphere-of-influence entry of
  /root/reference/raytracer/LimitedRelativisticRenderEngine.py:224,265 (flat ray_cast hit on the "isBH"
      sphere, minus the sphere centre).

`camera_rays(..., jitter="mt19937")` reproduces the reference's Mersenne-Twister stream bit for bit
(config 1, the parity bundle); `jitter="philox"` is the counter-based stream that the device-side
generator (csrc/raygen.cu) also produces, used for the full-size frames.
"""
from __future__ import annotations

import math
import random

import numpy as np

# BASELINE.json configs 1/2/4: camera off every axis, looking at the hole.
CFG_CAMERA_POS = (120.0, -80.0, 40.0)
CFG_FOV = 0.6
CFG_M = 1.0
CFG_R_SPHERE = 60.0
CFG_SEED = 42


def look_at_rotation(cam_pos, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
    """3x3 camera-to-world rotation; camera looks along its local -z, local +y is up (Blender convention)."""
    c = np.asarray(cam_pos, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - c
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, dtype=np.float64))
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    return np.stack([right, upv, -fwd], axis=1)


def euler_xyz_rotation(rx, ry, rz):
    """Blender 'XYZ' Euler -> matrix (what `matrix_world.to_euler()` + `Vector.rotate` apply, RRE:182,229)."""
    cx, sx = math.cos(rx), math.sin(rx)
    cy, sy = math.cos(ry), math.sin(ry)
    cz, sz = math.cos(rz), math.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85


def philox4x32_10(counter_lo: np.ndarray, seed: int):
    """Philox-4x32-10 on counters (ctr0=low 32 bits of index, ctr1=high 32 bits, 0, 0), key=(seed, 0).

    Returns four uint32 arrays.  csrc/raygen.cu implements the same function on the device; the two
    are compared bit for bit in the tests.
    """
    idx = np.asarray(counter_lo, dtype=np.uint64)
    c0 = idx & np.uint64(0xFFFFFFFF)
    c1 = idx >> np.uint64(32)
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def _u01_from_u32_pair(hi, lo):
    """53-bit uniform in [0,1) from two uint32 (same construction as MT's random())."""
    a = hi.astype(np.uint64) >> np.uint64(5)
    b = lo.astype(np.uint64) >> np.uint64(6)
    return (a * np.float64(67108864.0) + b) * (1.0 / 9007199254740992.0)


def camera_rays(width, height, samples=1, fov_x=CFG_FOV, fov_y=CFG_FOV, rotation=None, seed=CFG_SEED,
                jitter="mt19937", first_ray=0, n_rays=None):
    """Unit directions [N,3] in the reference's s -> y -> x order (N = samples*height*width).

    `first_ray`/`n_rays` select a contiguous slice of that order (used when sharding a frame).
    """
    if rotation is None:
        rotation = look_at_rotation(CFG_CAMERA_POS)
    n_total = samples * height * width
    if n_rays is None:
        n_rays = n_total - first_ray
    idx = np.arange(first_ray, first_ray + n_rays, dtype=np.int64)
    x = idx % width
    y = (idx // width) % height
    aspect = height / width
    dx = 1.0 / width
    dy = aspect / height
    if jitter == "mt19937":
        rng = random.Random()
        rng.seed(seed)
        u = np.array([rng.random() for _ in range(2 * (first_ray + n_rays))], dtype=np.float64)
        u = u[2 * first_ray:]
        u1, u2 = u[0::2], u[1::2]
    elif jitter == "philox":
        r0, r1, r2, r3 = philox4x32_10(idx, seed)
        u1 = _u01_from_u32_pair(r0, r1)
        u2 = _u01_from_u32_pair(r2, r3)
    elif jitter == "none":
        u1 = u2 = np.full(n_rays, 0.5)
    else:
        raise ValueError(f"unknown jitter {jitter!r}")
    xr = fov_x * (x - int(width / 2)) / width + dx * (u1 - 0.5)
    yr = fov_y * (y - int(height / 2)) / height * aspect + dy * (u2 - 0.5)
    local = np.stack([xr, yr, -np.ones_like(xr)], axis=1)
    d = local @ np.asarray(rotation, dtype=np.float64).T
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d


def sphere_entry(origin, directions, r_sphere, center=(0.0, 0.0, 0.0)):
    """First intersection of flat rays with the sphere of influence, relative to its centre.

    Returns (entry_pos[N,3], hit_mask[N]); rows with hit_mask False are NaN.  A camera INSIDE the sphere (where the
    RRE / CAM engines put it, RelativisticRenderEngine.py:278, RelativisticRenderEngineCamEdition.py:212) starts every
    ray at the camera itself.
    """
    o = np.asarray(origin, dtype=np.float64) - np.asarray(center, dtype=np.float64)
    d = np.asarray(directions, dtype=np.float64)
    if o @ o < r_sphere * r_sphere:
        return np.broadcast_to(o, d.shape).copy(), np.ones(d.shape[0], dtype=bool)
    od = d @ o
    disc = od * od - (o @ o - r_sphere * r_sphere)
    hit = (disc >= 0.0) & (od < 0.0)
    s = -od - np.sqrt(np.where(hit, disc, np.nan))
    hit &= s >= 0.0
    p = o[None, :] + s[:, None] * d
    p[~hit] = np.nan
    return p, hit


def config_bundle(width, height, samples=1, jitter="mt19937", cam_pos=CFG_CAMERA_POS, fov=CFG_FOV,
                  r_sphere=CFG_R_SPHERE, seed=CFG_SEED, first_ray=0, n_rays=None):
    """(entry_pos, entry_dir) for the BASELINE.json camera (configs 1, 2, 4). Every ray hits the sphere."""
    rot = look_at_rotation(cam_pos)
    d = camera_rays(width, height, samples, fov, fov, rot, seed, jitter, first_ray, n_rays)
    p, hit = sphere_entry(cam_pos, d, r_sphere)
    if not hit.all():
        raise ValueError("camera rays miss the sphere of influence")
    return p, d


def random_impact_bundle(n, cam_dist=200.0, fov=0.6, r_sphere=CFG_R_SPHERE, seed=CFG_SEED, width=1920, height=1080):
    """Config 3: 1920x1080 frame, camera outside the sphere on a generic axis; rays that miss are dropped."""
    cam = np.array([0.6, -0.64, 0.48]) * cam_dist
    rot = look_at_rotation(cam)
    d = camera_rays(width, height, 1, fov, fov, rot, seed, "philox", 0, None)
    p, hit = sphere_entry(cam, d, r_sphere)
    p, d = p[hit], d[hit]
    if n is not None:
        p, d = p[:n], d[:n]
    return p, d


def near_critical_bundle(n, M=CFG_M, r_sphere=CFG_R_SPHERE, b_lo=5.0, b_hi=5.4, seed=CFG_SEED, in_plane=False):
    """Config 5: conserved impact parameter b = L/E uniform in [b_lo, b_hi] M, entry on the sphere,
    orbital-plane orientation uniform random (or the equatorial plane when in_plane)."""
    rng = np.random.default_rng(seed)
    b = rng.uniform(b_lo, b_hi, n) * M
    rs = 2.0 * M
    # unit coordinate direction d at radius R with flat impact parameter b_flat:
    #   b = b_flat / sqrt(1 - (rs/R) (b_flat/R)^2)     ->  solve for b_flat
    R = r_sphere
    b_flat = b / np.sqrt(1.0 + (rs / R) * (b / R) ** 2)
    sin_a = b_flat / R                      # angle between -e_r and d
    cos_a = np.sqrt(1.0 - sin_a**2)
    if in_plane:
        e1 = np.tile([1.0, 0.0, 0.0], (n, 1))
        e2 = np.tile([0.0, 1.0, 0.0], (n, 1))
        ang = rng.uniform(0, 2 * np.pi, n)
        c, s = np.cos(ang)[:, None], np.sin(ang)[:, None]
        e1, e2 = c * e1 + s * e2, -s * e1 + c * e2
    else:
        e1 = rng.normal(size=(n, 3))
        e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
        v = rng.normal(size=(n, 3))
        v -= np.sum(v * e1, axis=1, keepdims=True) * e1
        e2 = v / np.linalg.norm(v, axis=1, keepdims=True)
    pos = R * e1
    d = -cos_a[:, None] * e1 + sin_a[:, None] * e2
    return pos, d, b


def conserved_impact_parameter(pos, d, M):
    """b = L/E for unit coordinate direction d at position pos."""
    pos = np.asarray(pos, dtype=np.float64)
    d = np.asarray(d, dtype=np.float64)
    r = np.linalg.norm(pos, axis=1)
    b_flat = np.linalg.norm(np.cross(pos, d), axis=1)
    return b_flat / np.sqrt(1.0 - (2.0 * M / r) * (b_flat / r) ** 2)
