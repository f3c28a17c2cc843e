"""CPU tests of the generated tableau products (csrc/rk45_tables.cuh): the Nystrom form used by the kernel must be
algebraically the same Dormand-Prince step as scipy's rk_step, error vector and dense output."""
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLES = os.path.join(ROOT, "blackhole_geodesic_calculator_b200", "csrc", "rk45_tables.cuh")


def load_tables():
    text = open(TABLES).read()
    vals = {}
    # tableau entries are written as "value,  // NAME = exact fraction"; the sincos constants (no " = ") are skipped
    for m in re.finditer(r"^\s+([-+0-9.e]+),\s+// (\w+) = ", text, flags=re.M):
        vals[m.group(2)] = float(m.group(1))
    return vals


def test_tables_regenerate_identically(tmp_path):
    before = open(TABLES).read()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "gen_tables.py")])
    assert open(TABLES).read() == before


def test_nystrom_form_equals_scipy_rk_step():
    from scipy.integrate._ivp.rk import RK45, rk_step
    T = load_tables()
    rng = np.random.default_rng(0)
    n = 3                                  # x, k in R^3; state y = (k, x), y' = (F(x, k), k)
    W1, W2 = rng.normal(size=(n, n)), rng.normal(size=(n, n))

    def F(x, k):
        return np.tanh(W1 @ x) + 0.3 * (W2 @ k) * k

    def fun(t, y):
        k, x = y[:n], y[n:]
        return np.concatenate([F(x, k), k])

    k0, x0, h = rng.normal(size=n), rng.normal(size=n), 0.37
    y0 = np.concatenate([k0, x0])
    K = np.empty((7, 2 * n))
    y_new, f_new = rk_step(fun, 0.0, y0, fun(0.0, y0), h, RK45.A, RK45.B, RK45.C, K)
    err = K.T @ RK45.E * h
    Q = K.T @ RK45.P

    # the kernel's arithmetic, in numpy, from the generated constants
    A = lambda j, l: T.get(f"A{j}{l}", 0.0)
    AA = lambda j, l: T.get(f"AA{j}{l}", 0.0)
    C = {1: 0.0, **{j: T[f"C{j}"] for j in range(2, 7)}}
    Ks = [F(x0, k0)]
    for j in range(2, 7):
        kt = k0 + h * sum(A(j, l) * Ks[l - 1] for l in range(1, j))
        xt = x0 + h * C[j] * k0 + h * h * sum(AA(j, l) * Ks[l - 1] for l in range(1, j - 1))
        Ks.append(F(xt, kt))
    B = {1: T["B1"], 2: 0.0, 3: T["B3"], 4: T["B4"], 5: T["B5"], 6: T["B6"]}
    kn = k0 + h * sum(B[l] * Ks[l - 1] for l in range(1, 7))
    xn = x0 + h * k0 + h * h * sum(T[f"BA{l}"] * Ks[l - 1] for l in range(1, 6))
    Ks.append(F(xn, kn))
    E = {1: T["E1"], 2: 0.0, 3: T["E3"], 4: T["E4"], 5: T["E5"], 6: T["E6"], 7: T["E7"]}
    ek = h * sum(E[l] * Ks[l - 1] for l in range(1, 8))
    ex = h * h * sum(T[f"EA{l}"] * Ks[l - 1] for l in range(1, 7))
    assert np.allclose(np.concatenate([kn, xn]), y_new, rtol=0, atol=1e-14)
    assert np.allclose(np.concatenate([ek, ex]), err, rtol=0, atol=1e-15)
    for j in range(7):  # stage values themselves
        assert np.allclose(Ks[j], K[j, :n], rtol=0, atol=1e-13)
    # dense output coefficients: momentum from P, position from PS / PA
    P = lambda j, c: (1.0 if (j == 1 and c == 0) else T.get(f"P{j}{c}", 0.0))
    for c in range(4):
        qk = sum(P(j, c) * Ks[j - 1] for j in range(1, 8))
        ps = 1.0 if c == 0 else T[f"PS{c}"]
        qx = k0 * ps + h * sum(T[f"PA{l}{c}"] * Ks[l - 1] for l in range(1, 7))
        assert np.allclose(qk, Q[:n, c], rtol=0, atol=1e-13)
        assert np.allclose(qx, Q[n:, c], rtol=0, atol=1e-13)


def test_sincos_table_scheme_of_the_attempt_loop():
    """The table-driven sincos of the RK45 attempt (csrc/geodesic_core.cuh sincos_lut), restated in numpy from the
    generated constants and table: angle = n pi/512 + r, two-term series, one rotation.  It must agree with the library
    sin / cos to < 2 ulp of the larger of the two over the angles an integration visits (and well beyond), and the two
    reduction constants written as literals in the kernel must be the table's and be encodable as FP64 immediates
    (zero low word)."""
    import struct
    text = open(TABLES).read()
    names = {m.group(1): int(m.group(2)) for m in re.finditer(r"^\s+(T_\w+) = (\d+),", text, flags=re.M)}
    body = text[text.index("__constant__ double c_tab"):]
    vals = [float(m.group(1)) for m in re.finditer(r"^\s+([-+0-9.e]+),\s+//", body[:body.index("};")], flags=re.M)]
    tab = lambda n: vals[names["T_" + n]]
    lut = np.array([[float(a), float(b)] for a, b in
                    re.findall(r"^\s+\{([-+0-9.e]+), ([-+0-9.e]+)\},", text[text.index("g_sincos_lut"):], flags=re.M)])
    assert lut.shape == (1024, 2)
    core = open(os.path.join(ROOT, "blackhole_geodesic_calculator_b200", "csrc", "geodesic_core.cuh")).read()
    assert "0x1.45f3p+7" in core and "0x1.921fbp-8" in core
    assert tab("L_N_OVER_PI") == float.fromhex("0x1.45f3p+7") and tab("L_P1") == float.fromhex("0x1.921fbp-8")
    for v in (tab("L_N_OVER_PI"), tab("L_P1")):
        assert struct.unpack("<Q", struct.pack("<d", v))[0] & 0xFFFFFFFF == 0
    import mpmath
    mpmath.mp.prec = 200
    assert abs(mpmath.mpf(tab("L_P1")) + mpmath.mpf(tab("L_P1T")) - mpmath.pi / 512) < mpmath.mpf(2) ** -80
    for i in range(0, 1024, 7):   # correctly rounded entries
        assert lut[i, 0] == float(mpmath.sin(i * mpmath.pi / 512)) or i % 512 == 0
        assert lut[i, 1] == float(mpmath.cos(i * mpmath.pi / 512)) or i % 512 == 256
    assert lut[0, 0] == 0.0 and lut[512, 0] == 0.0 and lut[256, 1] == 0.0 and lut[768, 1] == 0.0

    rng = np.random.default_rng(5)
    th = np.concatenate([rng.uniform(-40.0, 40.0, 200000), rng.uniform(0.0, np.pi, 200000), rng.uniform(-6000.0, 6000.0, 20000),
                         np.arange(-2048, 2048) * np.pi / 512, np.array([0.0, 1e-300, np.pi / 2, np.pi, 3.0, -3.0])])
    big = 6755399441055744.0
    tn = th * tab("L_N_OVER_PI") + big                  # the kernel's FMA; a plain product + sum differs only in which
    dn = tn - big                                       # neighbouring n is picked exactly at a rounding tie
    n = dn.astype(np.int64)
    r = (th - dn * tab("L_P1")) - dn * tab("L_P1T")     # dn * L_P1 is exact (21-bit constant): same value as the FMA chain
    # 512/pi is cut to 21 bits (relative error 3.2e-7): the cell index drifts by 0.3 cells at |theta| = 6000, i.e. |r|
    # may reach 0.81 cells there instead of 0.5 - the two-term series still have > 2 digits to spare at that size
    assert np.abs(r[np.abs(th) <= 40.0]).max() < np.pi / 1024 * 1.01 and np.abs(r).max() < 0.9 * np.pi / 512
    sc = lut[n & 1023]
    z = r * r
    ps = z * tab("LS2") + tab("LS1")
    pc = z * tab("LC2") - 0.5
    sr = r + r * (z * ps)
    cm = z * pc
    s = sc[:, 1] * sr + (sc[:, 0] * cm + sc[:, 0])
    c = -sc[:, 0] * sr + (sc[:, 1] * cm + sc[:, 1])
    ref_s = np.array([float(mpmath.sin(mpmath.mpf(float(t)))) for t in th[::97]])
    ref_c = np.array([float(mpmath.cos(mpmath.mpf(float(t)))) for t in th[::97]])
    ulp = np.spacing(np.maximum(np.abs(ref_s), np.abs(ref_c)))
    assert (np.abs(s[::97] - ref_s) / ulp).max() < 2.5 and (np.abs(c[::97] - ref_c) / ulp).max() < 2.5
    # against numpy everywhere (numpy's sin / cos are < 1 ulp themselves)
    ulp_all = np.spacing(np.maximum(np.abs(np.sin(th)), np.abs(np.cos(th))))
    assert (np.abs(s - np.sin(th)) / ulp_all).max() < 3.5 and (np.abs(c - np.cos(th)) / ulp_all).max() < 3.5
    assert np.abs(s * s + c * c - 1.0).max() < 1e-15
