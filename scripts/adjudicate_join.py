"""Outlier adjudication, join: three-way matrix GPU <-> scipy, port <-> scipy, GPU <-> port on the rays on which any two
of the three disagree beyond 1e-6 (config 5 in random planes: every ray of the 2^20; configs 3 and 2: the GPU <-> port
outliers plus a random control sample), with the per-ray conditioning of the reference's method
(oracle/port.conditioning: the oracle re-run with every RHS evaluation perturbed by one ulp).

Inputs: gpurun_out/adj_cfg5_3d_cpu.npz (scripts/adjudicate_cpu.py, here), gpurun_out/adj_*_gpu.npz
(scripts/adjudicate_gpu.py, B200 box).  Outputs: profiles/r2a_adjudication.json (the matrix) and
tests/golden/parity_outliers.npz (entry states + scipy / port / GPU results + conditioning of those rays)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blackhole_geodesic_calculator_b200 import raygen  # noqa: E402
from oracle import port, schwarzschild_ref as R  # noqa: E402

GO = os.path.join(ROOT, "gpurun_out")
SEEDS = (11, 23, 37, 41, 53, 67, 71, 83)
K_TOL = 10.0


def dev(ap, ad, bp, bd):
    with np.errstate(invalid="ignore"):
        v = np.maximum(np.abs(ap - bp).max(axis=1) / 60.0, np.abs(ad - bd).max(axis=1))
    return np.where(np.isfinite(v), v, np.inf)


def pair(name, a, b, sens, esc):
    """a, b: dict(pos, dir, status, attempt (or None), accept)"""
    d = dev(a["pos"], a["dir"], b["pos"], b["dir"])
    same = a["accept"] == b["accept"]
    if a.get("attempt") is not None and b.get("attempt") is not None:
        same &= a["attempt"] == b["attempt"]
    tol = np.maximum(1e-6, K_TOL * sens)
    return {
        "pair": name, "status_flips": int((a["status"] != b["status"]).sum()),
        "step_counts_differ": int((~same).sum()),
        "escaped_beyond_1e-6": int((esc & (d > 1e-6)).sum()),
        "of_which_same_step_counts": int((esc & same & (d > 1e-6)).sum()),
        "max_dev_same_steps": float(d[esc & same].max(initial=0.0)),
        "max_dev_differing_steps": float(d[esc & ~same].max(initial=0.0)),
        f"beyond_max(1e-6,{K_TOL:g}*conditioning)": int((esc & (d > tol)).sum()),
    }, d


def main():
    out = {"rule": f"a ray is held to max(1e-6, {K_TOL:g} x conditioning), conditioning = largest move of the oracle's "
                   f"own exit state over {len(SEEDS)} runs with every RHS evaluation perturbed by <= 1 ulp "
                   "(oracle/port.conditioning); statuses must be equal on every ray",
           "sets": {}}
    gold = {}
    # ---------------- config 5, random planes: all 2^20 rays have scipy + port; GPU on its outliers + control
    c = np.load(os.path.join(GO, "adj_cfg5_3d_cpu.npz"))
    g = np.load(os.path.join(GO, "adj_cfg5_3d_gpu.npz"))
    p, d, b = raygen.near_critical_bundle(1 << 20, in_plane=False)
    p, d = np.ascontiguousarray(p), np.ascontiguousarray(d)
    o = port.trace(p, d)
    assert np.array_equal(o["status"], c["port_status"])
    sens, moved = port.conditioning(p, d, base=o, seeds=SEEDS)
    sc = dict(pos=c["scipy_pos"], dir=c["scipy_dir"], status=c["scipy_status"], attempt=(c["scipy_nfev"] - 2) // 6,
              accept=c["scipy_accept"])
    po = dict(pos=o["exit_pos"], dir=o["exit_dir"], status=o["status"], attempt=o["n_attempt"], accept=o["n_accept"])
    esc = (sc["status"] == 0) & (po["status"] == 0)
    m_ps, d_ps = pair("port<->scipy (all 1048576 rays)", po, sc, sens, esc)
    gi = g["idx"]
    gp = dict(pos=g["gpu_pos"], dir=g["gpu_dir"], status=g["gpu_status"], attempt=g["gpu_attempt"], accept=g["gpu_accept"])
    sub = lambda D: {k: v[gi] for k, v in D.items()}
    m_gp, d_gp = pair("GPU<->port (all rays; the listed counts are exact, deviations from the saved outliers + control)",
                      gp, sub(po), sens[gi], esc[gi] & (gp["status"] == 0))
    m_gs, d_gs = pair("GPU<->scipy (on the GPU<->port outliers + 4096 control rays)", gp, sub(sc), sens[gi],
                      esc[gi] & (gp["status"] == 0))
    nrm = np.cross(p, d)
    nz = np.abs(nrm[:, 2]) / np.linalg.norm(nrm, axis=1)
    both = (d_gp > 1e-6) & esc[gi]
    out["sets"]["cfg5_random_planes"] = {
        "rays": int(len(p)), "escaped": int(esc.sum()), "conditioning_over_1e-7": int((sens > 1e-7).sum()),
        "rays_whose_steps_or_status_move_under_1ulp_jitter": int(moved.sum()),
        "matrix": [m_gs, m_ps, m_gp],
        "gpu_port_outliers_that_are_also_port_scipy_outliers": int((d_ps[gi][both] > 1e-6).sum()),
        "gpu_port_outliers": int(both.sum()),
        "median_ratio_gpu_port_over_port_scipy_on_them": float(np.median(d_gp[both] / np.maximum(d_ps[gi][both], 1e-300))),
        "control_sample_median_dev": {"gpu_port": float(np.median(d_gp[g["is_control"] & esc[gi]])),
                                      "port_scipy": float(np.median(d_ps[gi][g["is_control"] & esc[gi]])),
                                      "gpu_scipy": float(np.median(d_gs[g["is_control"] & esc[gi]]))},
        "outliers_by_pole_proximity(port<->scipy)": [
            {"nz_range": [lo, hi], "escaped": int((esc & (nz >= lo) & (nz < hi)).sum()),
             "beyond_1e-6": int((esc & (nz >= lo) & (nz < hi) & (d_ps > 1e-6)).sum())}
            for lo, hi in ((0, 1e-3), (1e-3, 1e-2), (1e-2, 1e-1), (1e-1, 1.0))],
    }
    # golden: union of every ray on which any pair disagrees (>1e-7) or steps differ, plus 512 control rays
    bad_ps = np.nonzero((d_ps > 1e-7) | (sc["accept"] != po["accept"]) | (sc["status"] != po["status"]))[0]
    ctrl = gi[g["is_control"]][:512]
    u5 = np.union1d(np.union1d(bad_ps, gi[~g["is_control"]]), ctrl)
    gold["cfg5_idx"] = u5.astype(np.int64)
    gold["cfg5_entry_pos"], gold["cfg5_entry_dir"] = p[u5], d[u5]
    for k, v in sc.items():
        gold["cfg5_scipy_" + k] = v[u5]
    for k, v in po.items():
        gold["cfg5_port_" + k] = v[u5]
    gold["cfg5_conditioning"] = sens[u5]
    # GPU results where we have them (NaN elsewhere)
    gpos = np.full((len(u5), 3), np.nan)
    gdir = np.full((len(u5), 3), np.nan)
    gst = np.full(len(u5), -1, np.int32)
    where = np.searchsorted(u5, gi)
    ok = (where < len(u5)) & (u5[np.minimum(where, len(u5) - 1)] == gi)
    gpos[where[ok]], gdir[where[ok]], gst[where[ok]] = g["gpu_pos"][ok], g["gpu_dir"][ok], g["gpu_status"][ok]
    gold["cfg5_gpu_r2a_pos"], gold["cfg5_gpu_r2a_dir"], gold["cfg5_gpu_r2a_status"] = gpos, gdir, gst

    # ---------------- configs 3 and 2: scipy on the GPU <-> port outliers + control
    R._build_rhs()
    for name in ("cfg3", "cfg2"):
        g = np.load(os.path.join(GO, f"adj_{name}_gpu.npz"))
        ep, ed = g["entry_pos"], g["entry_dir"]
        keep = ~g["is_control"]
        keep[np.nonzero(g["is_control"])[0][:1024]] = True   # all outliers + 1024 control rays
        ep, ed = np.ascontiguousarray(ep[keep]), np.ascontiguousarray(ed[keep])
        s_pos, s_dir, s_st, s_nfev, s_acc = R.trace_pool(ep, ed, max(1, (os.cpu_count() or 2) - 1), chunk=64)[:5]
        o = port.trace(ep, ed)
        sens, moved = port.conditioning(ep, ed, base=o, seeds=SEEDS)
        sc = dict(pos=s_pos, dir=s_dir, status=s_st, attempt=(s_nfev - 2) // 6, accept=s_acc)
        po = dict(pos=o["exit_pos"], dir=o["exit_dir"], status=o["status"], attempt=o["n_attempt"], accept=o["n_accept"])
        gp = dict(pos=g["gpu_pos"][keep], dir=g["gpu_dir"][keep], status=g["gpu_status"][keep],
                  attempt=g["gpu_attempt"][keep], accept=g["gpu_accept"][keep])
        esc = (sc["status"] == 0) & (po["status"] == 0) & (gp["status"] == 0)
        with open(os.path.join(GO, "adj_gpu_summary.json")) as f:
            full = json.load(f)[name]
        out["sets"][name] = {
            "rays_full_set": full["rays"], "gpu_port_full_set": full,
            "subset": f"{int((~g['is_control'][keep]).sum())} GPU<->port outliers (> 1e-7) + 1024 control rays",
            "matrix": [pair("GPU<->scipy", gp, sc, sens, esc)[0], pair("port<->scipy", po, sc, sens, esc)[0],
                       pair("GPU<->port", gp, po, sens, esc)[0]]}
        gold[name + "_idx"] = g["idx"][keep].astype(np.int64)
        gold[name + "_entry_pos"], gold[name + "_entry_dir"] = ep, ed
        for k, v in sc.items():
            gold[name + "_scipy_" + k] = v
        for k, v in po.items():
            gold[name + "_port_" + k] = v
        gold[name + "_conditioning"] = sens
        gold[name + "_gpu_r2a_pos"], gold[name + "_gpu_r2a_dir"], gold[name + "_gpu_r2a_status"] = gp["pos"], gp["dir"], gp["status"]
    gold["seeds"] = np.array(SEEDS)
    gold["k_tol"] = np.array(K_TOL)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "parity_outliers.npz"), **gold)
    with open(os.path.join(ROOT, "profiles", "r2a_adjudication.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
