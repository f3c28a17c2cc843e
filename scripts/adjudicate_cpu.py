"""Outlier adjudication, CPU half (runs in the build container, no GPU): the REAL scipy path
(oracle/schwarzschild_ref.py, solve_ivp RK45) and its C restatement (oracle/rk45_port.c) on every ray of
config 5 in random planes (2^20 rays, b in [5.0, 5.4] M).  Saves per-ray results of both so that the GPU half
(scripts/adjudicate_gpu.py, run on the B200 box) can be joined into the three-way matrix
GPU <-> scipy, port <-> scipy, GPU <-> port (scripts/adjudicate_join.py).

    python scripts/adjudicate_cpu.py [--set cfg5_3d|cfg3|cfg2_sample] [--procs P] [--out FILE]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blackhole_geodesic_calculator_b200 import raygen  # noqa: E402
from oracle import port, schwarzschild_ref as R  # noqa: E402


def rays_of(name):
    if name == "cfg5_3d":
        p, d, _ = raygen.near_critical_bundle(1 << 20, in_plane=False)
    elif name == "cfg3":
        p, d = raygen.random_impact_bundle(None)
    else:
        raise SystemExit("unknown set " + name)
    return np.ascontiguousarray(p), np.ascontiguousarray(d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="cfg5_3d")
    ap.add_argument("--procs", type=int, default=max(1, (os.cpu_count() or 2) - 2))
    ap.add_argument("--out", default=None)
    ap.add_argument("--indices", default=None, help=".npy of ray indices: run scipy on those only")
    a = ap.parse_args()
    out = a.out or os.path.join(ROOT, "gpurun_out", f"adj_{a.set}_cpu.npz")
    p, d = rays_of(a.set)
    idx = np.arange(p.shape[0]) if a.indices is None else np.load(a.indices)
    t0 = time.time()
    o = port.trace(p[idx], d[idx])
    t1 = time.time()
    print(f"port: {len(idx)} rays in {t1 - t0:.1f} s", flush=True)
    R._build_rhs()
    s_pos, s_dir, s_st, s_nfev, s_acc = R.trace_pool(p[idx], d[idx], a.procs, chunk=512)[:5]
    print(f"scipy: {len(idx)} rays in {time.time() - t1:.1f} s on {a.procs} processes", flush=True)
    np.savez_compressed(out, idx=idx, scipy_pos=s_pos, scipy_dir=s_dir, scipy_status=s_st, scipy_nfev=s_nfev,
                        scipy_accept=s_acc, port_pos=o["exit_pos"], port_dir=o["exit_dir"], port_status=o["status"],
                        port_nfev=o["nfev"], port_accept=o["n_accept"], port_attempt=o["n_attempt"])
    dev = np.maximum(np.abs(s_pos - o["exit_pos"]).max(axis=1) / 60.0, np.abs(s_dir - o["exit_dir"]).max(axis=1))
    esc = (s_st == 0) & (o["status"] == 0)
    same = (s_nfev == o["nfev"]) & (s_acc == o["n_accept"])
    print(f"status flips {int((s_st != o['status']).sum())}, step sequences differ on {int((~same).sum())} rays, "
          f"escaped beyond 1e-6: {int((dev[esc] > 1e-6).sum())} (of which same-steps {int((dev[esc & same] > 1e-6).sum())}), "
          f"max dev same-steps {dev[esc & same].max():.3e}, max dev differing {dev[esc & ~same].max(initial=0):.3e}")
    print("wrote", out)


if __name__ == "__main__":
    main()
