"""Host-side mirrors of the reference's call shapes for the geodesic path, on top of the batched `trace()`.

The reference engines reach the solver through three curvedpy interfaces (curvedpy itself is not part of the
reference repository); each class below keeps the names, argument meaning and result layout visible at the
reference's call sites, and adds a `*_batch` method that does a whole frame or tile in one GPU call:

  * `GeodesicIntegratorSchwarzschild.calc_trajectory`  <- raytracer/RelativisticRenderEngine.py:134,293-297,307-310
  * `SchwarzschildGeodesic.ray_trace` / `.approximateCurveEnd`
                                                       <- raytracer/LimitedRelativisticRenderEngine.py:90,273-279,308-319
  * `RelativisticCamera.run` / `.ray_blackhole_hit` / `.ray_end`
                                                       <- raytracer/RelativisticRenderEngineCamEdition.py:206-215,225-228

Trajectory polylines (what `checkHitDisk`, LimitedRelativisticRenderEngine.py:413-438, scans and what
`calc_trajectory(nr_points_curve=...)` returns) are sampled in flight on linspace(0, curve_end, N) up to the
termination time, like solve_ivp's t_eval does for the reference.
"""
from __future__ import annotations

import math

import numpy as np

from . import api, raygen


class GeodesicIntegratorSchwarzschild:
    """RRE call shape: fixed affine length, no sphere; end state = last sample (RRE.py:293-294,307-308)."""

    def __init__(self, mass=1.0, time_like=False, verbose=False, device=0, rtol=1e-3, atol=1e-6, eps_horizon=0.01):
        if time_like:
            raise NotImplementedError("only null geodesics (time_like=False) are on the reference's render path "
                                      "(RelativisticRenderEngine.py:134)")
        self.mass = float(mass)
        self.r_s = 2.0 * self.mass
        self.device = device
        self.rtol, self.atol, self.eps_horizon = rtol, atol, eps_horizon

    def calc_trajectories_batch(self, k0_xyz, x0_xyz, max_step=math.inf, curve_end=50.0, R_end=math.inf, mode="parity"):
        """k0_xyz, x0_xyz: [N,3].  Returns (end_dir[N,3] unit, end_loc[N,3], hit_blackhole[N] bool,
        start_inside_hole[N] bool, status[N])."""
        x0 = np.ascontiguousarray(x0_xyz, dtype=np.float64).reshape(-1, 3)
        k0 = np.ascontiguousarray(k0_xyz, dtype=np.float64).reshape(-1, 3)
        r_sphere = float(R_end)
        exit_pos, exit_dir, status = api.trace(x0, k0, self.mass, r_sphere, self.rtol, self.atol, max_step=max_step,
                                               eps_horizon=self.eps_horizon, lambda_max=float(curve_end), mode=mode,
                                               device=self.device)
        return exit_dir, exit_pos, status == api.CAPTURED, status == api.START_INSIDE_HOLE, status

    def calc_trajectory(self, k0_xyz, x0_xyz, max_step=math.inf, curve_end=50.0, nr_points_curve=50, verbose=False):
        """Per-ray drop-in: returns (k_xyz[3,N'], x_xyz[3,N'], result).  x_xyz holds the trajectory sampled on
        linspace(0, curve_end, nr_points_curve) up to the termination time (N' <= nr_points_curve), as curvedpy
        returns it (RelativisticRenderEngine.py:293-294,299-300); k_xyz holds the start direction in column 0 and
        the unit end direction in column -1 (the only column the engine reads, RRE.py:307-308), NaN in between."""
        k0 = np.asarray(k0_xyz, dtype=np.float64).reshape(1, 3)
        x0 = np.asarray(x0_xyz, dtype=np.float64).reshape(1, 3)
        ms = math.inf if (max_step is None or max_step == -1) else float(max_step)
        ep, ed, st, poly, cnt = api.trace(x0, k0, self.mass, math.inf, self.rtol, self.atol, max_step=ms,
                                          eps_horizon=self.eps_horizon, lambda_max=float(curve_end),
                                          device=self.device, polyline=max(2, int(nr_points_curve)))
        status = int(st[0])
        result = {"start_inside_hole": status == api.START_INSIDE_HOLE, "hit_blackhole": status == api.CAPTURED,
                  "status": status}
        c = max(int(cnt[0]), 1)
        x_xyz = poly[0, :c].T.copy()
        if status == api.START_INSIDE_HOLE:
            x_xyz = x0.T.copy()
        k_xyz = np.full_like(x_xyz, np.nan)
        k_xyz[:, 0] = k0[0]
        k_xyz[:, -1] = ed[0]
        return k_xyz, x_xyz, result


def spacetime_ray_cast_batch(integrator: GeodesicIntegratorSchwarzschild, origin, directions, bh_loc=(0.0, 0.0, 0.0),
                             max_step=math.inf, curve_end=50.0):
    """Batched body of `RelativisticRenderEngine.spacetime_ray_cast` (RRE.py:271-313) for one camera origin
    and N primary directions: returns (hit[N], hit_bh[N], end_dir[N,3], end_loc[N,3]) with `hit` all False
    exactly as the reference hard-wires it (RRE.py:305)."""
    d = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
    o = np.asarray(origin, dtype=np.float64) - np.asarray(bh_loc, dtype=np.float64)
    x0 = np.broadcast_to(o, d.shape).copy()
    end_dir, end_loc, hit_bh, inside, _ = integrator.calc_trajectories_batch(d, x0, max_step, curve_end)
    hit_bh = hit_bh | inside  # "Camera INSIDE blackhole" returns hit_bh=True (RRE.py:311-313)
    return np.zeros(d.shape[0], bool), hit_bh, end_dir, end_loc


class SchwarzschildGeodesic:
    """LIM call shape: entry point on the sphere of influence -> exit point/direction or capture
    (LimitedRelativisticRenderEngine.py:273-278).  The solver works in units of r_s (M = 1/2): the sphere
    object of radius |loc_hit| is `ratio_obj_to_blackhole` horizon radii large (LIM.py:488, README.md:60).

    `coordinates`: chart in which `loc_hit` / `direction` are read and `end_loc` / `end_dir` / the polyline returned.
    Default "isotropic": this curvedpy generation "uses the Schwarzschild metric in cartesian coordinates"
    (README.md:174) and README Fig. 5 / Fig. 6, drawn with it, are reproduced to the pixel in the isotropic chart
    and in no other (tests/test_readme_figures.py).  "schwarzschild" reads the same numbers as
    x = r sin(th) cos(ph) of the Schwarzschild radius (the newer generation's chart, RelativisticRenderEngine.py:289)."""

    def __init__(self, metric="schwarzschild", device=0, rtol=1e-3, atol=1e-6, eps_horizon=0.01,
                 coordinates="isotropic"):
        if metric != "schwarzschild":
            raise NotImplementedError(f"metric {metric!r}: only 'schwarzschild' is on the reference's render path")
        if coordinates not in api.COORDS:
            raise ValueError(f"coordinates must be one of {sorted(api.COORDS)}")
        self.metric = metric
        self.coordinates = coordinates
        self.device = device
        self.rtol, self.atol, self.eps_horizon = rtol, atol, eps_horizon

    @staticmethod
    def approximateCurveEnd(ratio_obj_to_blackhole):
        """Affine-length bound in r_s units.  curvedpy's own formula is not in the reference; its commented
        predecessor is `50 + 2*50*(ratio/20 - 1)` (LIM.py:279), which under-runs for small spheres, so the
        bound used is the larger of that and 10 sphere radii (a bound that is not reached changes nothing)."""
        r = float(ratio_obj_to_blackhole)
        return max(50.0 + 2.0 * 50.0 * (r / 20.0 - 1.0), 10.0 * r)

    def ray_trace_batch(self, directions, loc_hits, exit_tolerance=0.2, ratio_obj_to_blackhole=30.0, curve_end=None,
                        max_step=math.inf, mode="parity", disk=None):
        """directions, loc_hits: [N,3] in scene units, relative to the sphere centre.  Returns
        (end_loc[N,3] scene units, end_dir[N,3] unit, hit_blackhole[N], outside[N], status[N]).
        disk=(R_in, R_out) in r_s units (the engine passes disk_R_in * ratio, disk_R_out * ratio,
        LimitedRelativisticRenderEngine.py:284-285) appends disk_xy[N,2] in r_s units (NaN = no hit): the
        `loc` of checkHitDisk (LIM.py:413-438) found in flight."""
        d = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
        p = np.ascontiguousarray(loc_hits, dtype=np.float64).reshape(-1, 3)
        ratio = float(ratio_obj_to_blackhole)
        scale = ratio / np.linalg.norm(p, axis=1)          # scene units -> r_s units, per ray
        lam = self.approximateCurveEnd(ratio) if curve_end is None else float(curve_end)
        ms = math.inf if (max_step is None or max_step == -1) else float(max_step)
        res = api.trace(p * scale[:, None], d, 0.5, ratio, self.rtol, self.atol, max_step=ms,
                        eps_horizon=self.eps_horizon, lambda_max=lam, mode=mode, device=self.device, disk=disk,
                        coords=self.coordinates)
        exit_pos, exit_dir, status = res[:3]
        hit_bh = status == api.CAPTURED
        # 'Outside': the ray did not end on the sphere within exit_tolerance (LIM.py:311-314)
        off = np.abs(np.linalg.norm(exit_pos, axis=1) - ratio) > exit_tolerance
        outside = ~hit_bh & (off | (status != api.ESCAPED))
        if disk is not None:
            return exit_pos / scale[:, None], exit_dir, hit_bh, outside, status, res[-1]
        return exit_pos / scale[:, None], exit_dir, hit_bh, outside, status

    @staticmethod
    def disk_shading_inputs(disk_xy, R_in, R_out, disk_phase=0.0, disk_mean=0.2, disk_stddev=0.3, disk_intensity=1.0):
        """The per-hit arithmetic of checkHitDisk after the crossing is known (LIM.py:424-436): returns
        (hit[N], texture_x[N], texture_y[N], intensity[N]); the texture lookup itself stays in Blender."""
        xd, yd = disk_xy[:, 0], disk_xy[:, 1]
        hit = np.isfinite(xd)
        R = np.sqrt(xd * xd + yd * yd)
        with np.errstate(all="ignore"):
            scale = (R - R_in) / (R_out - R_in)
            intensity = disk_intensity * np.exp(-((scale - disk_mean) ** 2) / (2 * disk_stddev**2)) / np.sqrt(
                2 * np.pi * disk_stddev)
            texture_x = (disk_phase + np.arccos(xd / R) * (yd / np.abs(yd))) / np.pi
        return hit, texture_x, scale, intensity

    def ray_trace(self, direction, loc_hit, exit_tolerance=0.2, ratio_obj_to_blackhole=30.0, curve_end=None,
                  max_step=math.inf, warnings=False, nr_points_curve=256):
        """Per-ray drop-in: (x, y, z, end_loc, end_dir, mes).  x, y, z are the trajectory polyline in r_s units
        (the unit in which the engine's checkHitDisk compares radii with disk_R_in * ratio,
        LimitedRelativisticRenderEngine.py:283-285), sampled on linspace(0, curve_end, nr_points_curve) up to the
        exit / capture, with the exact end point appended."""
        d = np.asarray(direction, float).reshape(1, 3)
        p = np.asarray(loc_hit, float).reshape(1, 3)
        ratio = float(ratio_obj_to_blackhole)
        scale = ratio / np.linalg.norm(p[0])
        lam = self.approximateCurveEnd(ratio) if curve_end is None else float(curve_end)
        ms = math.inf if (max_step is None or max_step == -1) else float(max_step)
        ep, ed, st, poly, cnt = api.trace(p * scale, d, 0.5, ratio, self.rtol, self.atol, max_step=ms,
                                          eps_horizon=self.eps_horizon, lambda_max=lam, device=self.device,
                                          polyline=max(2, int(nr_points_curve)), coords=self.coordinates)
        status = int(st[0])
        hit_bh = status == api.CAPTURED
        off = abs(np.linalg.norm(ep[0]) - ratio) > exit_tolerance
        mes = {"hit_blackhole": hit_bh, "status": status}
        if not hit_bh and (off or status != api.ESCAPED):
            mes["error"] = "Outside"
        pts = np.vstack([poly[0, :int(cnt[0])], ep[0][None, :]])
        return pts[:, 0], pts[:, 1], pts[:, 2], ep[0] / scale, ed[0], mes


class Conversions:
    """Host-side coordinate conversions under curvedpy's name (used by the engines only for a debug print,
    RelativisticRenderEngine.py:289-291, RelativisticRenderEngineCamEdition.py:349): Cartesian position + tangent
    <-> spherical (r, theta, phi) position + coordinate-basis components, the same map the tracer applies on the
    device when a ray enters and leaves the integration (csrc/trace_kernel.cuh init_state / exit_state)."""

    @staticmethod
    def convert_xyz_to_sph(x_xyz, k_xyz):
        x, y, z = (float(v) for v in x_xyz)
        kx, ky, kz = (float(v) for v in k_xyz)
        r = math.sqrt(x * x + y * y + z * z)
        rho2 = x * x + y * y
        rho = math.sqrt(rho2)
        th, ph = math.acos(z / r), math.atan2(y, x)
        k_r = (x * kx + y * ky + z * kz) / r
        with np.errstate(all="ignore"):   # on the polar axis the spherical tangent is singular: inf / nan, no exception
            k_th = float(np.float64(z * (x * kx + y * ky) - rho2 * kz) / np.float64(r * r * rho))
            k_ph = float(np.float64(x * ky - y * kx) / np.float64(rho2))
        return np.array([r, th, ph]), np.array([k_r, k_th, k_ph])

    @staticmethod
    def convert_sph_to_xyz(x_sph, k_sph):
        r, th, ph = (float(v) for v in x_sph)
        k_r, k_th, k_ph = (float(v) for v in k_sph)
        st, ct, sp, cp = math.sin(th), math.cos(th), math.sin(ph), math.cos(ph)
        x = np.array([r * st * cp, r * st * sp, r * ct])
        k = np.array([k_r * st * cp + r * ct * cp * k_th - r * st * sp * k_ph,
                      k_r * st * sp + r * ct * sp * k_th + r * st * cp * k_ph,
                      k_r * ct - r * st * k_th])
        return x, k


class ApproxSchwarzschildGeodesic:
    """Call shape of curvedpy's tabulated approximate tracer (LimitedRelativisticRenderEngine.py:39-40,97-101,269):
    `ApproxSchwarzschildGeodesic(ratio_obj_to_blackhole=, exit_tolerance=).generatedRayTracer(loc, direction)` ->
    (end_loc, end_dir, mes).  The reference interpolates a pre-computed table because the exact solve is slow; here
    the exact solve is the fast path, so the same call is answered exactly (no table, no interpolation error) and
    the engine's `approx` branch keeps working unchanged.  `generatedRayTracer_batch` is the batched form."""

    def __init__(self, ratio_obj_to_blackhole=30.0, exit_tolerance=0.2, device=0, coordinates="isotropic"):
        self.ratio_obj_to_blackhole = float(ratio_obj_to_blackhole)   # attributes read back by the engine (LIM.py:97-98)
        self.exit_tolerance = float(exit_tolerance)
        self._exact = SchwarzschildGeodesic(device=device, coordinates=coordinates)

    def generatedRayTracer_batch(self, locs, directions):
        end_loc, end_dir, hit_bh, outside, status = self._exact.ray_trace_batch(
            directions, locs, exit_tolerance=self.exit_tolerance, ratio_obj_to_blackhole=self.ratio_obj_to_blackhole)
        return end_loc, end_dir, hit_bh, outside, status

    def generatedRayTracer(self, loc, direction):
        end_loc, end_dir, hit_bh, outside, status = self.generatedRayTracer_batch(
            np.asarray(loc, float).reshape(1, 3), np.asarray(direction, float).reshape(1, 3))
        mes = {"hit_blackhole": bool(hit_bh[0]), "status": int(status[0])}
        if outside[0]:
            mes["error"] = "Outside"
        return end_loc[0], end_dir[0], mes


class RelativisticCamera:
    """CAM call shape: a whole frame traced ahead of shading; the consumer reads `ray_blackhole_hit[iy,ix]`
    and `ray_end[iy,ix,3:6]` (RelativisticRenderEngineCamEdition.py:206-215,225-228).  `a` (spin) must be 0:
    only Schwarzschild is on the path."""

    def __init__(self, resolution=(64, 64), field_of_view=(0.6, 0.6), a=0.0, M=1.0, camera_location=raygen.CFG_CAMERA_POS,
                 camera_rotation_euler=None, r_sphere=raygen.CFG_R_SPHERE, samples=1, seed=raygen.CFG_SEED,
                 jitter="none", max_step=math.inf, verbose=False, device=0):
        if a != 0.0:
            raise NotImplementedError("Kerr (a != 0) is not on the reference's render path")
        self.resolution = (int(resolution[0]), int(resolution[1]))  # [height, width] (CamEdition.py:206)
        self.field_of_view = (float(field_of_view[0]), float(field_of_view[1]))
        self.M, self.r_sphere = float(M), float(r_sphere)
        self.camera_location = np.asarray(camera_location, dtype=np.float64)
        self.rotation = (raygen.look_at_rotation(self.camera_location) if camera_rotation_euler is None
                         else raygen.euler_xyz_rotation(*camera_rotation_euler))
        self.samples, self.seed, self.jitter, self.max_step, self.device = samples, seed, jitter, max_step, device
        self.ray_blackhole_hit = None
        self.ray_end = None
        self.ray_status = None

    def run(self, verbose=False, verbose_lvl=0, mode="parity"):
        h, w = self.resolution
        d = raygen.camera_rays(w, h, self.samples, self.field_of_view[0], self.field_of_view[1], self.rotation,
                               self.seed, self.jitter)
        p, hit = raygen.sphere_entry(self.camera_location, d, self.r_sphere)
        n = d.shape[0]
        exit_pos = np.full((n, 3), np.nan)
        exit_dir = d.copy()                      # rays that miss the sphere keep their flat direction
        status = np.full(n, api.MISSED_SPHERE, dtype=np.int32)  # never entered the curved region
        if hit.any():
            ep, ed, st = api.trace(p[hit], d[hit], self.M, self.r_sphere, max_step=self.max_step, mode=mode,
                                   device=self.device)
            exit_pos[hit], exit_dir[hit], status[hit] = ep, ed, st
        shape = (self.samples, h, w) if self.samples > 1 else (h, w)
        self.ray_status = status.reshape(shape)
        self.ray_blackhole_hit = (status == api.CAPTURED).astype(np.int64).reshape(shape)
        self.ray_end = np.concatenate([exit_pos, exit_dir], axis=1).reshape(shape + (6,))
        return self

    def save(self, path):
        np.savez_compressed(path, ray_blackhole_hit=self.ray_blackhole_hit, ray_end=self.ray_end,
                            ray_status=self.ray_status, resolution=self.resolution, field_of_view=self.field_of_view,
                            M=self.M, r_sphere=self.r_sphere, camera_location=self.camera_location)

    def load(self, path):
        z = np.load(path)
        self.ray_blackhole_hit, self.ray_end, self.ray_status = z["ray_blackhole_hit"], z["ray_end"], z["ray_status"]
        return self
