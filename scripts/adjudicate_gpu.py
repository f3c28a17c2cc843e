"""Outlier adjudication, GPU half (runs on the B200 box): the CUDA path against the C restatement on every ray of
config 5 in random planes (2^20), config 3 (1920x1080) and config 2 (1024^2 x 5 spp).  Writes, per set, the rays on
which the two disagree in any way (status, attempt/accept counts, or exit state beyond `--tol`) together with the GPU
results, plus a fixed random control sample, to gpurun_out/adj_<set>_gpu.npz.  The CPU half
(scripts/adjudicate_cpu.py) runs the REAL scipy path on the same generators; scripts/adjudicate_join.py builds the
three-way matrix.  Nothing here reads /root/reference."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blackhole_geodesic_calculator_b200 import api, raygen  # noqa: E402
from oracle import port  # noqa: E402


def rays_of(name):
    if name == "cfg5_3d":
        p, d, _ = raygen.near_critical_bundle(1 << 20, in_plane=False)
    elif name == "cfg3":
        p, d = raygen.random_impact_bundle(None)
    elif name == "cfg2":
        p, d = raygen.config_bundle(1024, 1024, 5, jitter="philox")
    else:
        raise SystemExit("unknown set " + name)
    return np.ascontiguousarray(p), np.ascontiguousarray(d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", default="cfg5_3d,cfg3,cfg2")
    ap.add_argument("--tol", type=float, default=1e-7)
    ap.add_argument("--control", type=int, default=4096)
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    summary = {}
    for name in a.sets.split(","):
        p, d = rays_of(name)
        ep, ed, st, cnt = api.trace(p, d, return_counters=True)
        o = port.trace(p, d)
        dev = np.maximum(np.abs(ep - o["exit_pos"]).max(axis=1) / 60.0, np.abs(ed - o["exit_dir"]).max(axis=1))
        dev = np.where(np.isfinite(dev), dev, np.inf)
        esc = (st == 0) & (o["status"] == 0)
        same = (cnt[0] == o["n_attempt"]) & (cnt[1] == o["n_accept"])
        bad = (st != o["status"]) | ~same | (esc & (dev > a.tol))
        rng = np.random.default_rng(7)
        ctrl = rng.choice(p.shape[0], size=min(a.control, p.shape[0]), replace=False)
        sel = np.union1d(np.nonzero(bad)[0], ctrl)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"adj_{name}_gpu.npz"), idx=sel, gpu_pos=ep[sel],
                            gpu_dir=ed[sel], gpu_status=st[sel], gpu_attempt=cnt[0][sel], gpu_accept=cnt[1][sel],
                            port_pos=o["exit_pos"][sel], port_dir=o["exit_dir"][sel], port_status=o["status"][sel],
                            port_attempt=o["n_attempt"][sel], port_accept=o["n_accept"][sel], entry_pos=p[sel],
                            entry_dir=d[sel], is_control=np.isin(sel, ctrl))
        summary[name] = dict(rays=int(p.shape[0]), escaped=int(esc.sum()), status_flips=int((st != o["status"]).sum()),
                             step_counts_differ=int((~same).sum()),
                             escaped_same_steps_beyond_1e6=int((esc & same & (dev > 1e-6)).sum()),
                             escaped_same_steps_max_dev=float(dev[esc & same].max(initial=0.0)),
                             escaped_diff_steps=int((esc & ~same).sum()),
                             escaped_diff_steps_beyond_1e6=int((esc & ~same & (dev > 1e-6)).sum()),
                             escaped_diff_steps_max_dev=float(dev[esc & ~same].max(initial=0.0)), saved=int(len(sel)))
    print(json.dumps(summary, indent=1))
    with open(os.path.join(ROOT, "gpurun_out", "adj_gpu_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
