"""ORACLE (test infrastructure): ctypes face of oracle/rk45_port.c — the scalar C restatement of
the reference path (curvedpy call sites RelativisticRenderEngine.py:293-294 /
LimitedRelativisticRenderEngine.py:273-278 over scipy RK45).  PARITY UNPINNED (see
oracle/schwarzschild_ref.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librk45_port.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "rk45_port.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        D = ctypes.c_double
        P = ctypes.c_void_p
        _lib.bhg_oracle_trace.argtypes = [P, P, ctypes.c_int64, D, D, D, D, D, D, D, ctypes.c_int, ctypes.c_int,
                                          P, P, P, P, P, P, P, D, D, P]
        _lib.bhg_oracle_trace.restype = ctypes.c_int
        _lib.bhg_oracle_max_threads.restype = ctypes.c_int
        _lib.bhg_oracle_set_jitter.argtypes = [ctypes.c_uint64]
        _lib.bhg_oracle_set_jitter.restype = None
    return _lib


def max_threads():
    return int(lib().bhg_oracle_max_threads())


def trace(entry_pos, entry_dir, M=1.0, r_sphere=60.0, rtol=1e-3, atol=1e-6, max_step=np.inf, eps_horizon=0.01,
          lambda_max=None, mode=0, nthreads=0, disk=None, jitter_seed=0):
    """Returns dict(exit_pos, exit_dir, status, nfev, n_accept, n_attempt, lam[, disk_xy]).
    `jitter_seed` != 0: conditioning probe, every RHS result is perturbed by -1/0/+1 ulp (see rk45_port.c)."""
    pos = np.ascontiguousarray(entry_pos, dtype=np.float64).reshape(-1, 3)
    dirs = np.ascontiguousarray(entry_dir, dtype=np.float64).reshape(-1, 3)
    n = pos.shape[0]
    if lambda_max is None:
        lambda_max = 10.0 * r_sphere
    out = dict(exit_pos=np.empty((n, 3)), exit_dir=np.empty((n, 3)), status=np.empty(n, np.int32),
               nfev=np.empty(n, np.int32), n_accept=np.empty(n, np.int32), n_attempt=np.empty(n, np.int32),
               lam=np.empty(n))
    if disk is not None:
        out["disk_xy"] = np.empty((n, 2))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib().bhg_oracle_set_jitter(int(jitter_seed))
    rc = lib().bhg_oracle_trace(p(pos), p(dirs), n, M, r_sphere, rtol, atol, max_step, eps_horizon, lambda_max,
                                int(mode), int(nthreads), p(out["exit_pos"]), p(out["exit_dir"]), p(out["status"]),
                                p(out["nfev"]), p(out["n_accept"]), p(out["n_attempt"]), p(out["lam"]),
                                float(disk[0]) if disk is not None else 0.0, float(disk[1]) if disk is not None else 0.0,
                                p(out["disk_xy"]) if disk is not None else None)
    lib().bhg_oracle_set_jitter(0)
    assert rc == 0
    return out


def conditioning(entry_pos, entry_dir, base=None, seeds=(11, 23, 37, 41), r_scale=None, **kw):
    """Per-ray conditioning of the reference's method: the largest change of the exit state (relative position,
    absolute direction) over `seeds` runs in which every RHS evaluation is perturbed by at most one ulp.  Two
    implementations of the same formulas differ by exactly that kind of rounding, so no parity bound below
    this number is meaningful for the ray.  Also returns whether the status or the step counts moved."""
    base = base or trace(entry_pos, entry_dir, **kw)
    r_scale = r_scale or (kw.get("r_sphere", 60.0) if np.isfinite(kw.get("r_sphere", 60.0)) else 1.0)
    n = base["status"].shape[0]
    sens = np.zeros(n)
    moved = np.zeros(n, dtype=bool)
    for s in seeds:
        q = trace(entry_pos, entry_dir, jitter_seed=int(s), **kw)
        with np.errstate(invalid="ignore"):
            dv = np.maximum(np.abs(q["exit_pos"] - base["exit_pos"]).max(axis=1) / r_scale,
                            np.abs(q["exit_dir"] - base["exit_dir"]).max(axis=1))
        sens = np.maximum(sens, np.where(np.isfinite(dv), dv, 0.0))
        moved |= (q["status"] != base["status"]) | (q["n_attempt"] != base["n_attempt"]) | (q["n_accept"] != base["n_accept"])
    return sens, moved
