"""ORACLE (test infrastructure, never shipped, never the thing measured as the product).

CPU restatement of the reference's per-ray Schwarzschild null-geodesic path.

PARITY UNPINNED (pinned only to README Fig. 5 / 6, tests/test_readme_figures.py): the reference
(`/root/reference`, bldevries/blackhole_geodesic_calculator) holds *no* tests or golden vectors for
this path - its only known answers are two figures of the README - and the arithmetic lives
in the third-party package `curvedpy` (un-vendored, un-pinned; README.md:24 names
"curvedpy v0.0.1"; not installed and not installable here).  This file therefore restates
curvedpy's *published* method and anchors it on the reference's call sites:

  * README.md:162-172  Schwarzschild metric in spherical coordinates (signature -+++, r_s)
  * README.md:133-135  Christoffel symbols  G^s_mn = 1/2 g^sr (d_m g_nr + d_n g_rm - d_r g_mn)
  * README.md:182      "The Christoffel Symbols are calculated using sympy"
  * README.md:198-209  first-order system  dk^a/dl = -G^a_nm k^m k^n ,  dx^b/dl = k^b  (8 eqs)
  * README.md:196,211  integrated with scipy.integrate.solve_ivp inside `calc_trajectory`
  * raytracer/RelativisticRenderEngine.py:134     ctor (mass, time_like=False)  -> null rays
  * raytracer/RelativisticRenderEngine.py:95      r_s = 2*M (geometrised units)
  * raytracer/RelativisticRenderEngine.py:289-291 Conversions().convert_xyz_to_sph(x, k)
  * raytracer/RelativisticRenderEngine.py:293-297 calc_trajectory(...) -> k_xyz, x_xyz, result;
                                                  result['start_inside_hole'], result['hit_blackhole']
  * raytracer/RelativisticRenderEngine.py:307-308 end state = last trajectory sample
  * raytracer/LimitedRelativisticRenderEngine.py:273-278,308-314 sphere entry -> exit contract,
                                                  mes['hit_blackhole'], mes['error']=='Outside'
  * no engine passes rtol/atol/method => scipy defaults RK45, rtol=1e-3, atol=1e-6
    (scipy 1.18.1, scipy/integrate/_ivp/rk.py:85-86).

The integrator is *the real scipy* `solve_ivp(method="RK45")`, so the step controller,
event location (brentq on the quartic dense output) and termination semantics are scipy's
own, not a re-derivation.  The scalar C port in `oracle/rk45_port.c` is validated against
this file.

Inferred (not evidenced by the reference; exposed as parameters): horizon event offset
`eps_horizon=0.01`, outer event `r - r_sphere` with direction=+1, state ordering
[k_t, t, k_r, r, k_th, th, k_ph, ph].
"""
from __future__ import annotations

import functools
import math

import numpy as np

# status codes (shared vocabulary with include/bhgeo.h)
ESCAPED = 0            # reached r_sphere going outward
CAPTURED = 1           # crossed r_s + eps_horizon  (reference: result['hit_blackhole'])
START_INSIDE_HOLE = 2  # r0 <= r_s + eps_horizon    (reference: result['start_inside_hole'])
LAMBDA_EXHAUSTED = 3   # affine length ran out      (reference LIM: mes['error']=='Outside'; RRE: normal end)
STEP_FAILED = 4        # scipy status -1 (step size too small) or non-finite state


@functools.lru_cache(maxsize=None)
def _build_rhs():
    """sympy: metric -> Christoffel -> geodesic RHS -> numpy callable (README.md:133-135,162-172,198-209)."""
    import sympy as sp

    t, r, th, ph, rs = sp.symbols("t r theta phi r_s", real=True)
    kt, kr, kth, kph = sp.symbols("k_t k_r k_theta k_phi", real=True)
    x = [t, r, th, ph]
    k = [kt, kr, kth, kph]
    g = sp.diag(-(1 - rs / r), 1 / (1 - rs / r), r**2, r**2 * sp.sin(th) ** 2)
    ginv = g.inv()
    dk = []
    for a in range(4):
        acc = 0
        for m in range(4):
            for n in range(4):
                gam = 0
                for p in range(4):
                    gam += sp.Rational(1, 2) * ginv[a, p] * (
                        sp.diff(g[n, p], x[m]) + sp.diff(g[p, m], x[n]) - sp.diff(g[m, n], x[p])
                    )
                acc += -gam * k[m] * k[n]
        dk.append(sp.simplify(acc))
    # state order [k_t, t, k_r, r, k_th, th, k_ph, ph]
    exprs = [dk[0], kt, dk[1], kr, dk[2], kth, dk[3], kph]
    f = sp.lambdify((kt, t, kr, r, kth, th, kph, ph, rs), exprs, modules="math", cse=True)
    # null condition g_mn k^m k^n = 0 solved for the future-directed k_t
    norm = sum(g[i, i] * k[i] ** 2 for i in range(4))
    kt_sol = [s for s in sp.solve(norm, kt)]
    kt_pos = sp.lambdify((kr, r, kth, th, kph, rs), kt_sol, modules="math")
    return f, kt_pos, [str(e) for e in exprs]


def rhs_expressions():
    """The simplified sympy expressions (for DESIGN.md / debugging)."""
    return _build_rhs()[2]


def rhs(y, rs):
    f = _build_rhs()[0]
    return np.array(f(*y, rs), dtype=np.float64)


def xyz_to_sph(x, k):
    """Position and coordinate tangent, Cartesian -> spherical (RelativisticRenderEngine.py:289-291)."""
    X, Y, Z = (float(v) for v in x)
    kx, ky, kz = (float(v) for v in k)
    rho2 = X * X + Y * Y
    r2 = rho2 + Z * Z
    r = math.sqrt(r2)
    rho = math.sqrt(rho2)
    th = math.acos(Z / r)
    ph = math.atan2(Y, X)
    xk = X * kx + Y * ky
    k_r = (xk + Z * kz) / r
    # on the polar axis (rho = 0) the spherical tangent is singular: IEEE semantics (inf/nan), as numpy
    # would give, instead of Python's ZeroDivisionError
    with np.errstate(all="ignore"):
        k_th = float(np.float64(Z * xk - rho2 * kz) / np.float64(r2 * rho))
        k_ph = float(np.float64(X * ky - Y * kx) / np.float64(rho2))
    return (r, th, ph), (k_r, k_th, k_ph)


def sph_to_xyz(xs, ks):
    r, th, ph = xs
    k_r, k_th, k_ph = ks
    st, ct = math.sin(th), math.cos(th)
    sp_, cp = math.sin(ph), math.cos(ph)
    X = r * st * cp
    Y = r * st * sp_
    Z = r * ct
    kx = k_r * st * cp + r * ct * cp * k_th - r * st * sp_ * k_ph
    ky = k_r * st * sp_ + r * ct * sp_ * k_th + r * st * cp * k_ph
    kz = k_r * ct - r * st * k_th
    return (X, Y, Z), (kx, ky, kz)


def null_kt(k_r, r, k_th, th, k_ph, rs):
    """Future-directed root of g_mn k^m k^n = 0 (time_like=False, RelativisticRenderEngine.py:134)."""
    s = math.sin(th)
    return r * math.sqrt(k_r * k_r + (r - rs) * r * (k_th * k_th + k_ph * k_ph * s * s)) / (r - rs)


def default_lambda_max(M, r_sphere):
    """Affine-length bound when the caller gives none: generous multiple of the sphere size.

    The reference's LIM engine uses `SW.approximateCurveEnd(ratio)` (curvedpy, absent;
    old heuristic LimitedRelativisticRenderEngine.py:279 `50 + 2*50*(ratio/20-1)` in r_s
    units = 200 M at ratio 30).  Any bound that is not reached gives identical results,
    so the default only has to be large: 10 * r_sphere.
    """
    return 10.0 * r_sphere


def trace_one(x0, k0, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, max_step=np.inf,
              eps_horizon=0.01, lambda_max=None, return_sol=False, disk=None, polyline=None):
    """One ray through real scipy solve_ivp.  Returns dict(exit_pos, exit_dir, status, nfev, n_accept, lam)."""
    from scipy.integrate import solve_ivp

    rs = 2.0 * M
    if lambda_max is None:
        lambda_max = default_lambda_max(M, r_sphere)
    x0 = np.asarray(x0, dtype=np.float64)
    k0 = np.asarray(k0, dtype=np.float64)
    r0 = float(np.sqrt(np.dot(x0, x0)))
    out = dict(exit_pos=np.full(3, np.nan), exit_dir=np.full(3, np.nan), status=START_INSIDE_HOLE,
               nfev=0, n_accept=0, lam=0.0, disk_xy=np.full(2, np.nan))
    if not (r0 > rs + eps_horizon):  # inside (or on) the horizon event surface: nothing to integrate
        return out
    (r, th, ph), (k_r, k_th, k_ph) = xyz_to_sph(x0, k0)
    try:
        k_t = null_kt(k_r, r, k_th, th, k_ph, rs)
    except (ValueError, OverflowError):
        k_t = math.nan
    y0 = np.array([k_t, 0.0, k_r, r, k_th, th, k_ph, ph])
    if not np.all(np.isfinite(y0)):
        # scipy refuses a non-finite initial state (base.py:21 ValueError); the batched contract reports
        # it per ray instead of raising: STEP_FAILED, no integration
        out["status"] = STEP_FAILED
        return out
    f = _build_rhs()[0]

    def fun(lam, y):
        try:
            return f(y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7], rs)
        except (ZeroDivisionError, ValueError, OverflowError):  # math-module form of an inf/nan RHS
            return [math.nan] * 8

    def hit_blackhole(lam, y):
        return y[3] - (rs + eps_horizon)

    hit_blackhole.terminal = True
    hit_blackhole.direction = 0
    events = [hit_blackhole]
    if np.isfinite(r_sphere):
        def reached_end(lam, y):
            return y[3] - r_sphere

        reached_end.terminal = True
        reached_end.direction = 1
        events.append(reached_end)

    if disk is not None:
        # equatorial-plane crossing z = r cos(theta) = 0 as a NON-terminal event on the same dense output: the
        # continuous form of checkHitDisk's polyline scan (LimitedRelativisticRenderEngine.py:413-438)
        def plane_crossing(lam, y):
            return math.cos(y[5])

        plane_crossing.terminal = False
        plane_crossing.direction = 0
        events.append(plane_crossing)

    t_eval = None
    if polyline is not None:
        # the reference asks for nr_points_curve samples on linspace(0, curve_end, N) (RelativisticRenderEngine.py:
        # 293-294: nr_points_curve=10000) and gets the ones up to the termination time back
        t_eval = np.linspace(0.0, lambda_max, int(polyline))
    with np.errstate(all="ignore"):
        res = solve_ivp(fun, (0.0, lambda_max), y0, method="RK45", events=events,
                        rtol=rtol, atol=atol, max_step=max_step, t_eval=t_eval)
    if polyline is not None:
        pts = np.full((int(polyline), 3), np.nan)
        cnt = res.y.shape[1] if res.y.ndim == 2 else 0
        for j in range(cnt):
            yj = res.y[:, j]
            (X, Y, Z), _ = sph_to_xyz((yj[3], yj[5], yj[7]), (yj[2], yj[4], yj[6]))
            pts[j] = (X, Y, Z)
        out["poly_xyz"], out["poly_count"] = pts, cnt
        # with t_eval, res.t / res.y hold the samples, not the terminal state: recover it from the events
        if res.status == 1:
            ev = 0 if len(res.t_events[0]) > 0 else 1
            res_t_end, res_y_end = res.t_events[ev][-1], res.y_events[ev][-1]
        else:
            from scipy.integrate import solve_ivp as _s
            r2 = _s(fun, (0.0, lambda_max), y0, method="RK45", events=events, rtol=rtol, atol=atol, max_step=max_step)
            res_t_end, res_y_end = r2.t[-1], r2.y[:, -1]
        res = _WithEnd(res, res_t_end, res_y_end)
    if disk is not None:
        r_in, r_out = disk
        for tc, yc in zip(res.t_events[-1], res.y_events[-1]):
            if r_in <= yc[3] <= r_out:   # first crossing inside the annulus (LIM.py:424)
                st = math.sin(yc[5])
                out["disk_xy"] = np.array([yc[3] * st * math.cos(yc[7]), yc[3] * st * math.sin(yc[7])])
                break
    yE = res.y[:, -1]
    lam = float(res.t[-1])
    if res.status == 1:
        # terminal event: scipy/integrate/_ivp/ivp.py:694-697 sets (t, y) to (root, sol(root))
        # before appending to ts/ys, so res.t[-1], res.y[:, -1] are the event state itself.
        if len(res.t_events[0]) > 0:
            status = CAPTURED
        else:
            status = ESCAPED
    elif res.status == 0:
        status = LAMBDA_EXHAUSTED
    else:
        status = STEP_FAILED
    if not np.all(np.isfinite(yE)):
        status = STEP_FAILED
    (X, Y, Z), (kx, ky, kz) = sph_to_xyz((yE[3], yE[5], yE[7]), (yE[2], yE[4], yE[6]))
    kn = math.sqrt(kx * kx + ky * ky + kz * kz)
    out.update(exit_pos=np.array([X, Y, Z]), exit_dir=np.array([kx, ky, kz]) / kn if kn > 0 else np.full(3, np.nan),
               status=status, nfev=int(res.nfev), n_accept=len(res.t) - 1, lam=lam, y_end=yE.copy())
    if return_sol:
        out["sol"] = res
    return out


class _WithEnd:
    """solve_ivp result whose .t / .y end with the terminal state (used when t_eval replaced them by samples)."""

    def __init__(self, res, t_end, y_end):
        self.__dict__.update(res)
        self.t = np.array([0.0, t_end])
        self.y = np.stack([res.y[:, 0] if res.y.size else y_end, y_end], axis=1)
        self._n_accept = None


def trace(entry_pos, entry_dir, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, max_step=np.inf,
          eps_horizon=0.01, lambda_max=None, disk=None, polyline=None):
    """Batched face of `trace_one` with the north-star signature (pure-Python loop; small N only)."""
    entry_pos = np.asarray(entry_pos, dtype=np.float64).reshape(-1, 3)
    entry_dir = np.asarray(entry_dir, dtype=np.float64).reshape(-1, 3)
    n = entry_pos.shape[0]
    exit_pos = np.empty((n, 3))
    exit_dir = np.empty((n, 3))
    status = np.empty(n, dtype=np.int32)
    nfev = np.empty(n, dtype=np.int32)
    n_accept = np.empty(n, dtype=np.int32)
    disk_xy = np.empty((n, 2))
    poly = np.full((n, int(polyline), 3), np.nan) if polyline is not None else None
    pcnt = np.zeros(n, dtype=np.int32)
    for i in range(n):
        o = trace_one(entry_pos[i], entry_dir[i], M, r_sphere, rtol, atol, max_step, eps_horizon, lambda_max,
                      disk=disk, polyline=polyline)
        exit_pos[i], exit_dir[i], status[i] = o["exit_pos"], o["exit_dir"], o["status"]
        nfev[i], n_accept[i] = o["nfev"], o["n_accept"]
        disk_xy[i] = o["disk_xy"]
        if polyline is not None and "poly_xyz" in o:
            poly[i], pcnt[i] = o["poly_xyz"], o["poly_count"]
    res = (exit_pos, exit_dir, status, nfev, n_accept)
    if disk is not None:
        res += (disk_xy,)
    if polyline is not None:
        res += (poly, pcnt)
    return res


def _pool_worker(args):
    pos, dirs, kw = args
    return trace(pos, dirs, **kw)


def trace_pool(entry_pos, entry_dir, processes, chunk=256, pool=None, **kw):
    """All-cores variant used for the CPU baseline (multiprocessing, chunks of >=256 rays)."""
    import multiprocessing as mp

    entry_pos = np.asarray(entry_pos, dtype=np.float64).reshape(-1, 3)
    entry_dir = np.asarray(entry_dir, dtype=np.float64).reshape(-1, 3)
    n = entry_pos.shape[0]
    jobs = [(entry_pos[i:i + chunk], entry_dir[i:i + chunk], kw) for i in range(0, n, chunk)]
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(processes)
    try:
        parts = pool.map(_pool_worker, jobs)
    finally:
        if own:
            pool.close()
            pool.join()
    return tuple(np.concatenate([p[j] for p in parts]) for j in range(len(parts[0])))
