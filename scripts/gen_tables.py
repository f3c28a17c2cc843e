#!/usr/bin/env python
"""Generates blackhole_geodesic_calculator_b200/csrc/rk45_tables.cuh: the Dormand-Prince 5(4) tableau of
scipy/_ivp/rk.py:538-566 and the products needed to run it in Nystrom form on a second-order system
(x' = k, k' = F(x, k)), computed in exact rational arithmetic and rounded once to double.

Nystrom form (algebraically identical to rk_step, rk.py:61-71, for the position half of the state):
  x_j   = x + h c_j k + h^2 sum_l AA[j][l] K_l          AA = A.A
  x_new = x + h k     + h^2 sum_l BA[l]    K_l          BA = B.A          (sum B = 1)
  err_x =               h^2 sum_l EA[l]    K_l          EA = E[:6].A + E[6] B   (sum E = 0)
  dense:  Qx_c = k PS[c] + h sum_l PA[l][c] K_l         PA = P[:6]^T.A + P[6] (x) B,  PS = column sums of P
where K_l are the stage values of the momentum derivative only.
"""
from fractions import Fraction as F
import os

C = [F(0), F(1, 5), F(3, 10), F(4, 5), F(8, 9), F(1)]
A = [[F(0)] * 5,
     [F(1, 5), 0, 0, 0, 0],
     [F(3, 40), F(9, 40), 0, 0, 0],
     [F(44, 45), F(-56, 15), F(32, 9), 0, 0],
     [F(19372, 6561), F(-25360, 2187), F(64448, 6561), F(-212, 729), 0],
     [F(9017, 3168), F(-355, 33), F(46732, 5247), F(49, 176), F(-5103, 18656)]]
A = [[F(v) for v in row] + [F(0)] for row in A]  # 6 x 6
B = [F(35, 384), F(0), F(500, 1113), F(125, 192), F(-2187, 6784), F(11, 84)]
E = [F(-71, 57600), F(0), F(71, 16695), F(-71, 1920), F(17253, 339200), F(-22, 525), F(1, 40)]
P = [[F(1), F(-8048581381, 2820520608), F(8663915743, 2820520608), F(-12715105075, 11282082432)],
     [F(0)] * 4,
     [F(0), F(131558114200, 32700410799), F(-68118460800, 10900136933), F(87487479700, 32700410799)],
     [F(0), F(-1754552775, 470086768), F(14199869525, 1410260304), F(-10690763975, 1880347072)],
     [F(0), F(127303824393, 49829197408), F(-318862633887, 49829197408), F(701980252875, 199316789632)],
     [F(0), F(-282668133, 205662961), F(2019193451, 616988883), F(-1453857185, 822651844)],
     [F(0), F(40617522, 29380423), F(-110615467, 29380423), F(69997945, 29380423)]]

assert sum(B) == 1 and sum(E) == 0
for j in range(6):
    assert sum(A[j]) == C[j]

AA = [[sum(A[j][i] * A[i][l] for i in range(6)) for l in range(6)] for j in range(6)]
BA = [sum(B[i] * A[i][l] for i in range(6)) for l in range(6)]
EA = [sum(E[j] * A[j][l] for j in range(6)) + E[6] * B[l] for l in range(6)]
PA = [[sum(P[j][c] * A[j][l] for j in range(6)) + P[6][c] * B[l] for c in range(4)] for l in range(6)]
PS = [sum(P[j][c] for j in range(7)) for c in range(4)]
for j in range(6):
    for l in range(6):
        assert AA[j][l] == 0 or l <= j - 2
assert BA[5] == 0 and PS[0] == 1


def d(x):
    return repr(float(x))


names = []
vals = []


def put(name, v):
    names.append(name)
    vals.append(v)


for j in range(1, 6):
    for l in range(j):
        put(f"A{j+1}{l+1}", A[j][l])
for j in range(1, 6):
    put(f"C{j+1}", C[j])
for j in range(2, 6):
    for l in range(j - 1):
        put(f"AA{j+1}{l+1}", AA[j][l])
for l in (0, 2, 3, 4, 5):
    put(f"B{l+1}", B[l])
for l in range(5):
    put(f"BA{l+1}", BA[l])
for l in (0, 2, 3, 4, 5, 6):
    put(f"E{l+1}", E[l])
for l in range(6):
    put(f"EA{l+1}", EA[l])
for l in (0, 2, 3, 4, 5, 6):
    for c in range(1, 4):
        put(f"P{l+1}{c}", P[l][c])
for c in range(1, 4):
    put(f"PS{c}", PS[c])
for l in range(6):
    for c in range(4):
        put(f"PA{l+1}{c}", PA[l][c])
# sincos: Cody-Waite pi/2 split and fdlibm kernel polynomials (k_sin.c / k_cos.c, public-domain coefficients)
extra = [
    ("TWO_OVER_PI", 6.36619772367581382433e-01),
    ("PIO2_1", float.fromhex("0x1.921fb54400000p+0")), ("PIO2_1T", 6.07710050650619224932e-11),
    ("PIO2_2", 6.07710050630396597660e-11), ("PIO2_2T", 2.02226624879595063154e-21),
    ("S1", -1.66666666666666324348e-01), ("S2", 8.33333333332248946124e-03), ("S3", -1.98412698298579493134e-04),
    ("S4", 2.75573137070700676789e-06), ("S5", -2.50507602534068634195e-08), ("S6", 1.58969099521155010221e-10),
    ("C1", 4.16666666666666019037e-02), ("C2", -1.38888888888741095749e-03), ("C3", 2.48015872894767294178e-05),
    ("C4", -2.75573143513906633035e-07), ("C5", 2.08757232129817482790e-09), ("C6", -1.13596475577881948265e-11),
]
# sincos for the hot loop: angle = n pi/512 + r, |r| <= pi/1024; (sin, cos)(n pi/512) from a 1024-entry table (correctly
# rounded, mpmath), sin r and cos r - 1 from two-term series (truncation r^7/5040 -> 1.7e-19 relative, r^6/720 -> 1.2e-18),
# then one rotation
import mpmath
mpmath.mp.prec = 200
_p64 = mpmath.pi / 512
# pi/512 and 512/pi cut to the 20 mantissa bits of a double's HIGH word: such constants are encodable as immediates of
# an FP64 instruction (no register read, no uniform register), n * L_P1 is exact for every
# n < 2^32, and the remainder L_P1T carries the other 53 bits
_l_p1 = float.fromhex("0x1.921fbp-8")
_l_n_over_pi = float.fromhex("0x1.45f3p+7")
extra += [
    ("L_N_OVER_PI", _l_n_over_pi), ("L_P1", _l_p1), ("L_P1T", float(_p64 - mpmath.mpf(_l_p1))),
    ("LS1", float(-mpmath.mpf(1) / 6)), ("LS2", float(mpmath.mpf(1) / 120)), ("LC2", float(mpmath.mpf(1) / 24)),
]
lut = [(float(mpmath.sin(n * _p64)), float(mpmath.cos(n * _p64))) for n in range(1024)]
for n in (0, 512):
    lut[n] = (0.0, lut[n][1])
for n in (256, 768):
    lut[n] = (lut[n][0], 0.0)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "blackhole_geodesic_calculator_b200", "csrc", "rk45_tables.cuh")
with open(out, "w") as f:
    f.write("// GENERATED by scripts/gen_tables.py (exact rational arithmetic, rounded once) - do not edit.\n")
    f.write("// Dormand-Prince 5(4) tableau (scipy/_ivp/rk.py:538-566) and its Nystrom-form products.\n")
    f.write("#pragma once\nnamespace bhg {\nnamespace tab {\n")
    f.write("enum : int {\n")
    for i, n in enumerate(names):
        f.write(f"    {n} = {i},\n")
    f.write(f"    N_RK = {len(names)},\n")
    for i, (n, _) in enumerate(extra):
        f.write(f"    T_{n} = {len(names) + i},\n")
    f.write(f"    N_ALL = {len(names) + len(extra)}\n}};\n")
    f.write("}  // namespace tab\n")
    f.write("__constant__ double c_tab[tab::N_ALL] = {\n")
    for n, v in zip(names, vals):
        f.write(f"    {d(v)},  // {n} = {v}\n")
    for n, v in extra:
        f.write(f"    {v!r},  // {n}\n")
    f.write("};\n")
    f.write("// (sin, cos)(n pi / 512), n = 0 .. 1023, correctly rounded; copied to shared memory by the trace kernel\n")
    f.write("__device__ const double g_sincos_lut[1024][2] = {\n")
    for sv, cv in lut:
        f.write(f"    {{{sv!r}, {cv!r}}},\n")
    f.write("};\n}  // namespace bhg\n")
print("wrote", out, len(names) + len(extra), "constants")
