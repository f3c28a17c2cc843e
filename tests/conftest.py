import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# stated parity tolerances (BASELINE.json north_star): 1e-6 relative in position (relative to the sphere
# radius / scene scale), 1e-6 rad in direction; status bit-exact outside |b - 3*sqrt(3) M| <= 1e-2 M
POS_RTOL = 1e-6
DIR_ATOL = 1e-6
B_CRIT_BAND = 1e-2


# Per-ray parity rule for full-size sets (profiles/r2a_adjudication.json).  The conditioning of
# a ray is how far the ORACLE's own exit state moves when every RHS evaluation is perturbed by <= 1 ulp
# (oracle/port.conditioning, 8 seeds).  Measured on all 2^20 rays of config 5 in random planes, for all three pairs
# GPU<->scipy, port<->scipy, GPU<->port alike:
#   * conditioning < 1e-7 (99.6 % of that set, all but a handful of rays of configs 2 and 3): identical step counts
#     and agreement within 1e-6 - no exceptions;
#   * otherwise the ray is ill-conditioned IN THE REFERENCE'S METHOD (pole-grazing and/or near-critical): statuses
#     still equal, deviation <= max(1e-6, 10 x conditioning); the conditioning is the maximum of 8 draws of a
#     heavy-tailed response, so at most 1 ray in 10^5 may exceed even that.
COND_SEEDS = (11, 23, 37, 41, 53, 67, 71, 83)
COND_WELL = 1e-7
COND_K = 10.0
COND_MAX_VIOLATION_FRAC = 1e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_usable():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests skip (instead of failing) on a machine without a CUDA device."""
    if _cuda_usable():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the product path has no CPU fallback (run with -m gpu on a B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def golden_kwargs(g):
    """Solver keyword arguments stored with a golden file."""
    kw = dict(M=float(g["M"]), r_sphere=float(g["r_sphere"]), rtol=float(g["rtol"]), atol=float(g["atol"]))
    if "lambda_max" in g:
        kw["lambda_max"] = float(g["lambda_max"])
    if "max_step" in g:
        kw["max_step"] = float(g["max_step"])
    return kw


def direction_angle(a, b):
    """Angle between unit vectors, atan2(|a x b|, a . b): an antiparallel (sign-flipped) direction gives pi, not 0."""
    a, b = np.asarray(a), np.asarray(b)
    return np.arctan2(np.linalg.norm(np.cross(a, b), axis=1), np.sum(a * b, axis=1))


def ray_deviation(got_pos, got_dir, ref_pos, ref_dir, scale):
    """max(relative position difference, direction angle) per ray; non-finite -> inf."""
    with np.errstate(invalid="ignore"):
        v = np.maximum(np.abs(got_pos - ref_pos).max(axis=1) / scale, direction_angle(got_dir, ref_dir))
    return np.where(np.isfinite(v), v, np.inf)


def assert_conditioned_parity(got_pos, got_dir, got_status, got_attempt, got_accept, ref, cond, scale, label=""):
    """The per-ray rule stated above.  `ref`: dict(exit_pos, exit_dir, status, n_attempt (or None), n_accept);
    `cond`: per-ray conditioning (oracle/port.conditioning).  Returns a dict of what was measured."""
    got_status, ref_status = np.asarray(got_status), np.asarray(ref["status"])
    assert np.array_equal(got_status, ref_status), \
        f"{label}: status differs on rays {np.nonzero(got_status != ref_status)[0][:10]}"
    cmp = np.isin(ref_status, (0, 3))
    dev = ray_deviation(got_pos, got_dir, ref["exit_pos"], ref["exit_dir"], scale)
    same = np.asarray(got_accept) == np.asarray(ref["n_accept"])
    if ref.get("n_attempt") is not None and got_attempt is not None:
        same &= np.asarray(got_attempt) == np.asarray(ref["n_attempt"])
    integ = ref_status != 2
    well = cond < COND_WELL
    bad_steps = integ & well & ~same
    assert not bad_steps.any(), f"{label}: well-conditioned rays with different step counts: {np.nonzero(bad_steps)[0][:10]}"
    bad_well = cmp & well & (dev > POS_RTOL)
    assert not bad_well.any(), \
        f"{label}: well-conditioned rays beyond 1e-6: {np.nonzero(bad_well)[0][:10]} dev {dev[bad_well][:10]}"
    viol = cmp & ~well & (dev > np.maximum(POS_RTOL, COND_K * cond))
    allowed = int(np.floor(COND_MAX_VIOLATION_FRAC * max(int(cmp.sum()), 1)))
    assert viol.sum() <= allowed, \
        f"{label}: {int(viol.sum())} ill-conditioned rays beyond max(1e-6, {COND_K:g} x conditioning) (allowed {allowed}): " \
        f"{np.nonzero(viol)[0][:10]} dev {dev[viol][:10]} cond {cond[viol][:10]}"
    out = dict(rays=int(len(ref_status)), compared=int(cmp.sum()), ill_conditioned=int((cmp & ~well).sum()),
               steps_differ=int((integ & ~same).sum()), beyond_1e6=int((cmp & (dev > POS_RTOL)).sum()),
               max_dev_well=float(dev[cmp & well].max(initial=0.0)), violations=int(viol.sum()), allowed=allowed)
    print(f"{label}: {out}")
    return out


def assert_parity(got_pos, got_dir, got_status, ref_pos, ref_dir, ref_status, scale, exclude=None,
                  pos_rtol=POS_RTOL, dir_atol=DIR_ATOL, captured_tol=1e-3):
    """Status bit-exact; exit position/direction within the stated tolerance for rays that left the sphere or
    used up their affine length — the states the reference consumes (LIM.py:317-319, RRE.py:246).  The exit
    state of a CAPTURED ray is never read by the reference (black pixel, LIM.py:308-309, RRE.py:242-244) and,
    for rays that wind many times around the photon sphere before falling in, round-off is amplified beyond any
    fixed bound (scipy vs its own C restatement differ by 1e-5 rad on such a ray), so it is only checked
    loosely.  `exclude` masks rays inside the stated critical band."""
    got_status = np.asarray(got_status)
    ref_status = np.asarray(ref_status)
    keep = np.ones(len(ref_status), bool) if exclude is None else ~exclude
    bad = keep & (got_status != ref_status)
    assert not bad.any(), f"status mismatch on rays {np.nonzero(bad)[0][:10]}: {got_status[bad][:10]} vs {ref_status[bad][:10]}"
    cmp = keep & np.isin(ref_status, (0, 3))
    dpos = np.abs(got_pos[cmp] - ref_pos[cmp]).max(initial=0.0) / scale
    ddir = direction_angle(got_dir[cmp], ref_dir[cmp]).max(initial=0.0)
    assert dpos <= pos_rtol, f"exit position differs by {dpos:.3e} (relative to {scale})"
    assert ddir <= dir_atol, f"exit direction differs by {ddir:.3e} rad"
    cap = keep & (ref_status == 1)
    if cap.any():
        cpos = np.abs(got_pos[cap] - ref_pos[cap]).max() / scale
        cdir = direction_angle(got_dir[cap], ref_dir[cap]).max()
        assert cpos <= captured_tol and cdir <= captured_tol, f"captured-ray end state differs by {cpos:.2e} / {cdir:.2e}"
    return dpos, ddir
