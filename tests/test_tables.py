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
