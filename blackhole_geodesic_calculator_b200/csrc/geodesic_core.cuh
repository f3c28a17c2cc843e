// Per-ray FP64 arithmetic of the Schwarzschild null-geodesic path, written for sm_100a.
//
// What it replaces (reference = /root/reference, bldevries/blackhole_geodesic_calculator):
//   curvedpy's calc_trajectory / ray_trace as called at raytracer/RelativisticRenderEngine.py:293-294
//   and raytracer/LimitedRelativisticRenderEngine.py:273-278, i.e. the geodesic RHS of README.md:198-209
//   (metric README.md:162-172, Christoffels README.md:133-135) under scipy's RK45 (README.md:196).
// The step controller, tolerances, initial step and event semantics follow scipy 1.18.1
// (_ivp/rk.py:8-11,14-71,85-176,538-566,715-738; _ivp/common.py:63-134; _ivp/ivp.py:52-158,659-699).
//
// B200-first organisation of the same discrete algorithm:
//  * The 8-variable system is x' = k, k' = F(x, k).  Only the four momentum derivatives K_j = F(stage j)
//    are kept per stage; the position halves of the stage sums, of y_new, of the error estimate and of the
//    dense output are formed from K_j with the pre-multiplied tableau products A.A, B.A, E.A, P.A
//    (rk45_tables.cuh, exact rational arithmetic).  This is algebraically the same Dormand-Prince step as
//    scipy's rk_step — same stages, same y_new, same error vector — but needs 28 instead of 56 doubles of
//    stage storage, which is what lets >= 4 warps per scheduler stay resident on the FP64 pipe.
//  * t and phi never enter the RHS, so their stage values are not formed at all (only t_new, phi_new and
//    their error terms are).
//  * One Newton-refined reciprocal per RHS, FMA-form sums, inverse tenth root by Newton instead of pow,
//    sincos with table constants; all constants come from __constant__ memory (uniform loads).
// Nothing here is translated from the reference (it contains no solver code); the CPU restatement of the
// reference path lives in oracle/.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "rk45_tables.cuh"

namespace bhg {

enum Status : int {
    ESCAPED = 0,
    CAPTURED = 1,
    START_INSIDE_HOLE = 2,
    LAMBDA_EXHAUSTED = 3,
    STEP_FAILED = 4,
};

// lane states of the warp work queue (negative: not a final status)
enum LaneState : int {
    LANE_RUNNING = -1,
    LANE_EMPTY = -4,
};

#define TAB(name) (c_tab[tab::name])

// ---------------------------------------------------------------------------------------------
// small FP64 building blocks
// ---------------------------------------------------------------------------------------------

// Comparisons of NON-NEGATIVE doubles on the integer pipe: for such values the order of the numbers is the order
// of their bit patterns.  The unsigned compare sends a NaN of either sign (and any negative value) to "huge",
// i.e. lt_nn(NaN, x) is false like the IEEE compare.  (Every FP64-pipe instruction saved is ~0.16 % of the
// kernel: profiles/r1m_experiments.txt.)
__device__ __forceinline__ bool lt_nn(double a, double b) {
    return (unsigned long long)__double_as_longlong(a) < (unsigned long long)__double_as_longlong(b);
}
__device__ __forceinline__ bool le_nn(double a, double b) {
    return (unsigned long long)__double_as_longlong(a) <= (unsigned long long)__double_as_longlong(b);
}
// |a| < c for c > 0, on the integer pipe (written on the 32-bit halves so it is not folded back into an FP64 |x|)
__device__ __forceinline__ bool abs_lt(double a, double c) {
    const unsigned long long v = ((unsigned long long)(unsigned)(__double2hiint(a) & 0x7fffffff) << 32) |
                                 (unsigned)__double2loint(a);
    return v < (unsigned long long)__double_as_longlong(c);
}
__device__ __forceinline__ double min_nn(double a, double b) { return lt_nn(b, a) ? b : a; }  // NaN a -> b
__device__ __forceinline__ double max_nn(double a, double b) { return lt_nn(a, b) ? b : a; }

// 1/a for normal, finite a (call sites guarantee that, or produce NaN/inf that the step controller
// rejects): MUFU.RCP64H seed, one third-order and one second-order refinement (5 DFMA).  Measured on
// B200 against IEEE division: bit-identical on the self-test sweep (profiles/r1b_selftest.txt).
__device__ __forceinline__ double fast_rcp5(double a) {
    double x0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(a));
    double e = fma(-a, x0, 1.0);
    double e2 = fma(e, e, e);
    double x1 = fma(x0, e2, x0);
    double e3 = fma(-a, x1, 1.0);
    return fma(x1, e3, x1);
}

// 3-DFMA variant (seed + one third-order step): measured max relative error 2.2e-16 (1 ulp) on B200.
// This is the reciprocal of the hot loop.
__device__ __forceinline__ double fast_rcp(double a) {
    double x0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(a));
    double e = fma(-a, x0, 1.0);
    double e2 = fma(e, e, e);
    return fma(x0, e2, x0);
}

// a^(-1/10) for a in [1e-12, 1e8]: float seed x0 (relative error e0 ~ 1e-6), then with d = a x0^10 - 1
// (|d| ~ 10 e0) the exact answer is x0 (1 + d)^(-1/10) = x0 (1 - d/10 + 11 d^2/200 - 77 d^3/2000 + ...); the
// truncation error 0.03 d^4 is < 1e-18 for |d| < 1e-4.  Dependency depth 9 instead of 14 for two Newton steps
// (this sits on the serial path between two attempts).  One Newton step follows only if the seed was poor.
__device__ __noinline__ double pow_cold(double a, double e) { return pow(a, e); }  // cold fallback, one copy

// 2^(e * log2(a)) in FP32 on the MUFU unit (lg2.approx / ex2.approx: relative error ~1e-6), as a double
__device__ __forceinline__ double pow_seed(double a, float e) {
    float l, r;
    const float af = (float)a;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(af));
    l *= e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l));
    return (double)r;
}

template <bool CHECK = true>
__device__ __forceinline__ double inv_tenth_root(double a) {
    const double x = pow_seed(a, -0.1f);
    const double x2 = x * x;
    const double x4 = x2 * x2;
    const double x8 = x4 * x4;
    const double d = fma(a, x8 * x2, -1.0);
    const double p = d * fma(d, fma(d, -0.0385, 0.055), -0.1);
    double res = fma(x, p, x);
    // seed worse than 1e-5 (never with the MUFU seed) or NaN: cold library fallback.  Tested after the result is
    // formed so that the compare overlaps the polynomial instead of sitting on the serial path.
    // CHECK = false: the step controller clamps its argument to [1e-11, 1e9] (NaN included), where the MUFU seed is
    // always within 1e-6 - the fallback would be dead code that still costs ten ALU instructions and a call frame
    if (CHECK && !abs_lt(d, 1e-4)) res = pow_cold(a, -0.1);
    return res;
}

// max(|a|, |b|): one FP64 compare with |.| modifiers plus two selects (a 7-instruction integer-pipe version
// measured the same, profiles/r1m_experiments.txt).
__device__ __forceinline__ double abs_max(double a, double b) { return fmax(fabs(a), fabs(b)); }

// The operand of larger magnitude, sign kept (the caller applies |.| as an operand modifier of its FMA): one compare and
// one 64-bit select.  fmax()'s NaN handling costs five more ALU instructions per call, each of which needs two register
// reads and therefore a cycle of its own (profiles/r2e_regread2.txt).  A NaN b is returned (NaN scale -> the step is
// rejected); a is an accepted state and never NaN.
__device__ __forceinline__ double larger_mag(double a, double b) { return fabs(a) > fabs(b) ? a : b; }

// a^(1/5) for a in [1e-30, 1e30] (initial-step heuristic): Newton on x^-5 = a for the inverse root, then
// a * x^4.  Float seed 1e-5 -> two steps -> 1e-18.
__device__ __forceinline__ double fifth_root(double a) {
    double x = pow_seed(a, -0.2f);
#pragma unroll
    for (int i = 0; i < 2; i++) {
        double x2 = x * x;
        double x5 = x2 * x2 * x;
        x = x * fma(-a, x5, 6.0) * 0.2;
    }
    double x2 = x * x;
    return a * (x2 * x2);
}

// sin and cos, < 1 ulp each: Cody-Waite reduction by pi/2 in two parts and the fdlibm kernel polynomials on
// [-pi/4, pi/4].  Branch-free.  The reduction is exact for |n| < 2^20 and stays within the argument's own
// ulp up to |th| ~ 1e9; beyond that (a state that has already blown up) the result is NaN, which the step
// controller treats like any other non-finite error norm.
template <bool GUARD = true>
__device__ __forceinline__ void sincos_tab(double th, double* s, double* c) {
    // n = rint(th * 2/pi) by the 1.5 * 2^52 trick: one FMA + one add, no FP64<->int conversion instructions
    const double big = 6755399441055744.0;
    const double tn = fma(th, TAB(T_TWO_OVER_PI), big);
    const int n = __double2loint(tn);
    const double dn = tn - big;
    double r = fma(-dn, TAB(T_PIO2_1), th);
    r = fma(-dn, TAB(T_PIO2_1T), r);
    // |th| >= 1e9 (0x41CDCD65 in the high word), inf or NaN -> NaN; integer compare keeps it off the FP64 pipe
    // (GUARD = false: the RK45 attempt tests theta_new once per attempt instead - angle_in_range() - which spares the
    // hot loop four ALU instructions per evaluation)
    if (GUARD) {
        const bool ok = (unsigned)(__double2hiint(th) & 0x7fffffff) < 0x41CDCD65u;
        r = ok ? r : __longlong_as_double(0x7ff8000000000000LL);
    }
    const double z = r * r;
    double ps = fma(z, TAB(T_S6), TAB(T_S5));
    double pc = fma(z, TAB(T_C6), TAB(T_C5));
    ps = fma(z, ps, TAB(T_S4));
    pc = fma(z, pc, TAB(T_C4));
    ps = fma(z, ps, TAB(T_S3));
    pc = fma(z, pc, TAB(T_C3));
    ps = fma(z, ps, TAB(T_S2));
    pc = fma(z, pc, TAB(T_C2));
    ps = fma(z, ps, TAB(T_S1));
    pc = fma(z, pc, TAB(T_C1));
    const double sr = fma(r, z * ps, r);                 // sin(r) = r + r (z S(z)): two distinct register operands
    const double cr = fma(z, fma(z, pc, -0.5), 1.0);     // cos(r) = 1 - z/2 + z^2 C(z)
    const double a = (n & 1) ? cr : sr;
    const double b = (n & 1) ? sr : cr;
    // quadrant signs by flipping the sign bit with integer ops (keeps the negations off the FP64 pipe)
    *s = __longlong_as_double(__double_as_longlong(a) ^ ((long long)(n & 2) << 62));
    *c = __longlong_as_double(__double_as_longlong(b) ^ ((long long)((n + 1) & 2) << 62));
}

// sin and cos for the RK45 attempt: th = n pi/512 + r with |r| <= pi/1024; (sin, cos)(n pi/512) come from a 1024-entry
// table in shared memory (16 KB), sin r and cos r - 1 from two-term series, and one rotation combines them.  14 FP64
// instructions and 3 ALU ones instead of 20 and 10 for sincos_tab (whose quadrant swap / sign logic costs a cycle per
// two-operand select, profiles/r2e_regread2.txt) and 6 constants instead of 15.  Error < 2 ulp.  No range guard: see
// angle_in_range().
constexpr int SINCOS_LUT_N = 1024;
__device__ __forceinline__ void sincos_lut(const double2* __restrict__ lut, double th, double* s, double* c) {
    // 512/pi and pi/512 cut to the high word of a double (scripts/gen_tables.py: same values as T_L_N_OVER_PI, T_L_P1):
    // immediates of the FMA, so the reduction reads no constant from a register
    const double big = 6755399441055744.0;
    const double tn = fma(th, 0x1.45f3p+7, big);
    const int n = __double2loint(tn);
    const double dn = tn - big;
    double r = fma(-dn, 0x1.921fbp-8, th);
    r = fma(-dn, TAB(T_L_P1T), r);
    const double2 sc = lut[n & (SINCOS_LUT_N - 1)];
    const double z = r * r;
    const double ps = fma(z, TAB(T_LS2), TAB(T_LS1));
    const double pc = fma(z, TAB(T_LC2), -0.5);
    const double sr = fma(r, z * ps, r);  // sin r
    const double cm = z * pc;             // cos r - 1
    *s = fma(sc.y, sr, fma(sc.x, cm, sc.x));
    *c = fma(-sc.x, sr, fma(sc.y, cm, sc.y));
}

// |th| < 1e9: the range in which sincos_tab's reduction is valid (integer compare on the high word; false for NaN)
__device__ __forceinline__ bool angle_in_range(double th) {
    return (unsigned)(__double2hiint(th) & 0x7fffffff) < 0x41CDCD65u;
}

// ---------------------------------------------------------------------------------------------
// right-hand sides: momentum derivatives only (positions' derivatives are the momenta themselves).
//   NK = 4 (parity): k = (k_t, k_r, k_th, k_ph), x = (t, r, th, ph)
//   NK = 3 (plane) : k = (k_t, k_r, k_ph),       x = (t, r, ph)      (theta = pi/2, k_th = 0)
// ---------------------------------------------------------------------------------------------
template <int NK>
struct Rhs;

template <>
struct Rhs<4> {
    // which position components the RHS reads (r and theta): only those get stage values
    __device__ __forceinline__ static constexpr bool needs_x(int i) { return i == 1 || i == 2; }
    __device__ __forceinline__ static void eval(const double (&k)[4], const double (&x)[4], double rs,
                                                double (&f)[4], const double2* __restrict__ lut = nullptr) {
        const double kt = k[0], kr = k[1], kth = k[2], kph = k[3], r = x[1], th = x[2];
        double s, c;
        if (lut) sincos_lut(lut, th, &s, &c);
        else sincos_tab<false>(th, &s, &c);
        const double rm = r - rs;
        const double p = r * rm;
        // one reciprocal for everything (two independent ones measured no faster on B200: r1 profiles)
        const double inv = fast_rcp(p * s);   // 1 / (r (r - rs) sin th)
        const double i_s = inv * p;           // 1 / sin th
        const double i_rrm = inv * s;         // 1 / (r (r - rs))
        const double i_r = i_rrm * rm;        // 1 / r
        const double hA = (0.5 * rs) * i_rrm; // rs / (2 r (r - rs))   (0.5 rs is loop-invariant)
        // Forms chosen for the register-read budget (profiles/r2e_regread2.txt): an FP64 instruction costs
        // max(2, distinct register operands) cycles, so squares (one operand) and products with constants are cheap and
        // three-variable FMAs are dear.
        const double u = (rm * i_r) * kt;             // (1 - rs/r) k_t
        const double d = fma(-u, u, kr * kr);         // k_r^2 - (1 - rs/r)^2 k_t^2
        const double kphs = kph * s;
        const double ang = fma(kphs, kphs, kth * kth);  // k_th^2 + k_ph^2 sin^2
        const double kr_r = kr * i_r;
        f[0] = ((-2.0 * hA) * kr) * kt;
        f[1] = fma(hA, d, rm * ang);
        f[2] = fma(kr_r * kth, -2.0, (kphs * kph) * c);
        f[3] = (-2.0 * kph) * fma(kth, c * i_s, kr_r);
    }
};

template <>
struct Rhs<3> {
    __device__ __forceinline__ static constexpr bool needs_x(int i) { return i == 1; }
    __device__ __forceinline__ static void eval(const double (&k)[3], const double (&x)[3], double rs,
                                                double (&f)[3], const double2* __restrict__ = nullptr) {
        const double kt = k[0], kr = k[1], kph = k[2], r = x[1];
        const double rm = r - rs;
        const double i_rrm = fast_rcp(r * rm);
        const double i_r = i_rrm * rm;
        const double hA = (0.5 * rs) * i_rrm;
        const double q = rm * i_r;
        const double w = (hA * q) * q;
        const double hAkr = hA * kr;
        f[0] = (-2.0 * hAkr) * kt;
        f[1] = fma(hAkr, kr, fma(-w * kt, kt, rm * (kph * kph)));
        f[2] = (-2.0 * kph) * (kr * i_r);
    }
};

// ---------------------------------------------------------------------------------------------
// one RK45 attempt in Nystrom form (rk_step + error estimate, scipy/_ivp/rk.py:14-71,105-109,143-146).
// On entry K[0] = F(k, x).  Fills K[1..6], kn, xn; returns sum_i (err_i / scale_i)^2 over all 2 NK
// components of the state.
// ---------------------------------------------------------------------------------------------
template <int NK>
__device__ __forceinline__ double knew_component(double ki, double K0, double K2, double K3, double K4, double K5,
                                                 double h) {
    double acc = TAB(B1) * K0;
    acc = fma(TAB(B3), K2, acc);
    acc = fma(TAB(B4), K3, acc);
    acc = fma(TAB(B5), K4, acc);
    acc = fma(TAB(B6), K5, acc);
    return fma(h, acc, ki);
}

template <int NK>
__device__ __forceinline__ double xnew_component(double xi, double ki, double K0, double K1, double K2, double K3,
                                                 double K4, double h, double h2) {
    double acc = TAB(BA1) * K0;
    acc = fma(TAB(BA2), K1, acc);
    acc = fma(TAB(BA3), K2, acc);
    acc = fma(TAB(BA4), K3, acc);
    acc = fma(TAB(BA5), K4, acc);
    return fma(h2, acc, fma(h, ki, xi));
}

template <int NK>
__device__ __forceinline__ double rk45_attempt(const double (&k)[NK], const double (&x)[NK], double (&K)[7][NK],
                                               double (&kn)[NK], double (&xn)[NK], const double h, const double rs,
                                               const double atol_over_rtol, const double inv_rtol2,
                                               const double2* __restrict__ lut = nullptr) {
    const double h2 = h * h;
    double kt[NK], xt[NK], hk[NK];
#pragma unroll
    for (int i = 0; i < NK; i++) {
        xt[i] = x[i];
        hk[i] = Rhs<NK>::needs_x(i) ? h * k[i] : 0.0;  // h k for the position stage values (r, theta only)
    }
    // ---- stage 2
    {
        const double ha = h * TAB(A21);
#pragma unroll
        for (int i = 0; i < NK; i++) {
            kt[i] = fma(ha, K[0][i], k[i]);
            if (Rhs<NK>::needs_x(i)) xt[i] = fma(TAB(C2), hk[i], x[i]);
        }
        Rhs<NK>::eval(kt, xt, rs, K[1], lut);
    }
    // ---- stage 3
    {
#pragma unroll
        for (int i = 0; i < NK; i++) {
            kt[i] = fma(h, fma(TAB(A32), K[1][i], TAB(A31) * K[0][i]), k[i]);
            if (Rhs<NK>::needs_x(i)) xt[i] = fma(h2, TAB(AA31) * K[0][i], fma(TAB(C3), hk[i], x[i]));
        }
        Rhs<NK>::eval(kt, xt, rs, K[2], lut);
    }
    // ---- stage 4
    {
#pragma unroll
        for (int i = 0; i < NK; i++) {
            kt[i] = fma(h, fma(TAB(A43), K[2][i], fma(TAB(A42), K[1][i], TAB(A41) * K[0][i])), k[i]);
            if (Rhs<NK>::needs_x(i))
                xt[i] = fma(h2, fma(TAB(AA42), K[1][i], TAB(AA41) * K[0][i]), fma(TAB(C4), hk[i], x[i]));
        }
        Rhs<NK>::eval(kt, xt, rs, K[3], lut);
    }
    // ---- stage 5
    {
#pragma unroll
        for (int i = 0; i < NK; i++) {
            kt[i] = fma(h, fma(TAB(A54), K[3][i], fma(TAB(A53), K[2][i], fma(TAB(A52), K[1][i], TAB(A51) * K[0][i]))),
                        k[i]);
            if (Rhs<NK>::needs_x(i))
                xt[i] = fma(h2, fma(TAB(AA53), K[2][i], fma(TAB(AA52), K[1][i], TAB(AA51) * K[0][i])),
                            fma(TAB(C5), hk[i], x[i]));
        }
        Rhs<NK>::eval(kt, xt, rs, K[4], lut);
    }
    // ---- stage 6 (c6 = 1)
    {
#pragma unroll
        for (int i = 0; i < NK; i++) {
            kt[i] = fma(h,
                        fma(TAB(A65), K[4][i],
                            fma(TAB(A64), K[3][i], fma(TAB(A63), K[2][i], fma(TAB(A62), K[1][i], TAB(A61) * K[0][i])))),
                        k[i]);
            if (Rhs<NK>::needs_x(i))
                xt[i] = fma(h2,
                            fma(TAB(AA64), K[3][i], fma(TAB(AA63), K[2][i], fma(TAB(AA62), K[1][i], TAB(AA61) * K[0][i]))),
                            x[i] + hk[i]);
        }
        Rhs<NK>::eval(kt, xt, rs, K[5], lut);
    }
    // ---- new state and FSAL stage
#pragma unroll
    for (int i = 0; i < NK; i++) {
        kn[i] = knew_component<NK>(k[i], K[0][i], K[2][i], K[3][i], K[4][i], K[5][i], h);
        xn[i] = xnew_component<NK>(x[i], k[i], K[0][i], K[1][i], K[2][i], K[3][i], K[4][i], h, h2);
    }
    Rhs<NK>::eval(kn, xn, rs, K[6], lut);
    // ---- error estimate, scaled (rk.py:143-146, common.py:63-65)
    double ek[NK], ex[NK], sk[NK], sx[NK];
#pragma unroll
    for (int i = 0; i < NK; i++) {
        double e = TAB(E1) * K[0][i];
        e = fma(TAB(E3), K[2][i], e);
        e = fma(TAB(E4), K[3][i], e);
        e = fma(TAB(E5), K[4][i], e);
        e = fma(TAB(E6), K[5][i], e);
        ek[i] = fma(TAB(E7), K[6][i], e);          // error / h   (momentum half)
        double g = TAB(EA1) * K[0][i];
        g = fma(TAB(EA2), K[1][i], g);
        g = fma(TAB(EA3), K[2][i], g);
        g = fma(TAB(EA4), K[3][i], g);
        g = fma(TAB(EA5), K[4][i], g);
        ex[i] = fma(TAB(EA6), K[5][i], g);         // error / h^2 (position half)
        // scale / rtol = max(|y|, |y_new|) + atol / rtol: an addition with one constant instead of an FMA with two
        // (the common factor 1 / rtol^2 of the squared norm is applied once, below)
        sk[i] = fabs(larger_mag(k[i], kn[i])) + atol_over_rtol;
        sx[i] = fabs(larger_mag(x[i], xn[i])) + atol_over_rtol;
    }
    // one reciprocal per group of four scales (two momentum/position pairs); scales are >= atol so the
    // products neither overflow nor underflow.  sum (e_i/scale_i)^2 = h^2 (S_k + h^2 S_x): h is applied once.
    double sumk = 0.0, sumx = 0.0;
#pragma unroll
    for (int i = 0; i + 1 < NK; i += 2) {
        const double pa = sk[i] * sx[i], pb = sk[i + 1] * sx[i + 1];
        const double inv = fast_rcp(pa * pb);
        const double ia = inv * pb, ib = inv * pa;  // 1/pa, 1/pb
        const double q0 = ek[i] * (ia * sx[i]), q1 = ex[i] * (ia * sk[i]);
        const double q2 = ek[i + 1] * (ib * sx[i + 1]), q3 = ex[i + 1] * (ib * sk[i + 1]);
        sumk = fma(q0, q0, sumk);
        sumx = fma(q1, q1, sumx);
        sumk = fma(q2, q2, sumk);
        sumx = fma(q3, q3, sumx);
    }
    if (NK & 1) {
        constexpr int i = NK - 1;
        const double inv = fast_rcp(sk[i] * sx[i]);
        const double q0 = ek[i] * (inv * sx[i]), q1 = ex[i] * (inv * sk[i]);
        sumk = fma(q0, q0, sumk);
        sumx = fma(q1, q1, sumx);
    }
    double esum = (h2 * inv_rtol2) * fma(h2, sumx, sumk);
    // a stage angle outside sincos_tab's range made this attempt meaningless: NaN rejects it (RHS evaluations are unguarded)
    if (NK == 4 && !angle_in_range(xn[NK == 4 ? 2 : 0])) esum = __longlong_as_double(0x7ff8000000000000LL);
    return esum;
}

// step-size factors 0.9 * err^(-1/5) (scipy/_ivp/rk.py:148-163) as a function of the SUM of squares
// esum = n * err^2 (n = 2 NK components): 0.9 (esum/n)^(-1/10) = (0.9 n^(1/10)) esum^(-1/10), so the division by
// n costs nothing.  Branch-free clamps: the caller already knows whether the step was accepted (esum < n) or
// rejected (esum >= n or NaN).
// step_factor_raw: f = 0.9 err^(-1/5) with esum clamped to [1e-11, 1e9] (NaN -> 1e9): f(1e-11) > 11 is above every
// acceptance cap, f(1e9) = 0.14 below the rejection floor, so the clamp never changes the final factor; the caller
// applies min(f, 10 or 1) on acceptance and max(f, 0.2) on rejection.
template <int N2>
__device__ __forceinline__ double step_factor_raw(double esum) {
    constexpr double c = (N2 == 8) ? 0.9 * 1.2311444133449163 : 0.9 * 1.1962311988513155;  // 0.9 * N2^(1/10)
    return c * inv_tenth_root<false>(min_nn(max_nn(esum, 1e-11), 1e9));
}
//   accepted: min(hi, f), hi = 10 (or 1 after a rejection); esum -> 0 gives hi (f(1e-11) > 11)
template <int N2>
__device__ __forceinline__ double step_factor_accept(double esum, double hi) {
    constexpr double c = (N2 == 8) ? 0.9 * 1.2311444133449163 : 0.9 * 1.1962311988513155;  // 0.9 * N2^(1/10)
    return min_nn(c * inv_tenth_root<false>(min_nn(max_nn(esum, 1e-11), 1e9)), hi);
}
//   rejected: max(0.2, f); huge / inf / NaN error norms give 0.2 (min_nn drops the NaN; f(1e9) = 0.14)
template <int N2>
__device__ __forceinline__ double step_factor_reject(double esum) {
    constexpr double c = (N2 == 8) ? 0.9 * 1.2311444133449163 : 0.9 * 1.1962311988513155;
    return max_nn(c * inv_tenth_root<false>(max_nn(min_nn(esum, 1e9), 1e-11)), 0.2);
}

// 10 * |nextafter(t, +inf) - t|  for t >= 0 (scipy/_ivp/rk.py:119)
__device__ __forceinline__ double min_step_at(double t) {
    double up = __longlong_as_double(__double_as_longlong(t) + 1);
    return 10.0 * (up - t);
}

// ---------------------------------------------------------------------------------------------
// Hairer's initial step (scipy/_ivp/common.py:109-134, order = 4, direction = +1); one RHS evaluation.
// f0 = (K0, k) in the (momentum, position) split.
// ---------------------------------------------------------------------------------------------
template <int NK>
__device__ __forceinline__ double initial_step(const double (&k)[NK], const double (&x)[NK], const double (&K0)[NK],
                                               double rs, double rtol, double atol, double interval,
                                               double max_step) {
    if (interval == 0.0) return 0.0;
    // scale_i = atol + |y_i| rtol; one reciprocal per (momentum, position) pair
    double isk[NK], isx[NK];
    double D0 = 0.0, D1 = 0.0;  // sums of squares: d0^2 = D0 / (2 NK), d1^2 = D1 / (2 NK)
#pragma unroll
    for (int i = 0; i < NK; i++) {
        const double sk = fma(fabs(k[i]), rtol, atol), sx = fma(fabs(x[i]), rtol, atol);
        const double inv = fast_rcp5(sk * sx);
        isk[i] = inv * sx;
        isx[i] = inv * sk;
        const double a = k[i] * isk[i], b = x[i] * isx[i];
        const double c = K0[i] * isk[i], d = k[i] * isx[i];
        D0 = fma(a, a, fma(b, b, D0));
        D1 = fma(c, c, fma(d, d, D1));
    }
    const double n2 = 2.0 * NK;
    // h0 = 0.01 d0 / d1 unless d0 < 1e-5 or d1 < 1e-5  (common.py:117-120), on the squares
    double h0 = (D0 < 1e-10 * n2 || D1 < 1e-10 * n2) ? 1e-6 : 0.01 * sqrt(D0 * fast_rcp5(D1));
    h0 = fmin(h0, interval);
    double k1[NK], x1[NK], F1[NK];
#pragma unroll
    for (int i = 0; i < NK; i++) {
        k1[i] = fma(h0, K0[i], k[i]);
        x1[i] = fma(h0, k[i], x[i]);
    }
    Rhs<NK>::eval(k1, x1, rs, F1);
    double D2 = 0.0;
#pragma unroll
    for (int i = 0; i < NK; i++) {
        const double a = (F1[i] - K0[i]) * isk[i];
        const double b = (k1[i] - k[i]) * isx[i];
        D2 = fma(a, a, fma(b, b, D2));
    }
    // d1^2 and d2^2 = D2 / (2 NK h0^2); h1 = (0.01 / max(d1, d2))^(1/5) = (max(d1,d2)^2 * 1e4)^(-1/10)
    const double ih0 = fast_rcp5(h0);
    const double d1sq = D1 / n2, d2sq = (D2 / n2) * (ih0 * ih0);
    const double msq = fmax(d1sq, d2sq);
    double h1;
    if (!(msq > 1e-30)) h1 = fmax(1e-6, h0 * 1e-3);  // d1 <= 1e-15 and d2 <= 1e-15
    else {
        const double arg = msq * 1e4;
        h1 = (arg > 1e-30 && arg < 1e30) ? inv_tenth_root(arg) : pow(arg, -0.1);
    }
    return fmin(fmin(100.0 * h0, h1), fmin(interval, max_step));
}

// ---------------------------------------------------------------------------------------------
// dense output (scipy/_ivp/rk.py:554-566,715-738): y(t_old + s h) = y_old + h * sum_c Q_c s^(c+1)
//   momentum: Q_c = sum_j K_j P[j][c]
//   position: Q_c = k PS[c] + h sum_l K_l PA[l][c]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dense_coeffs_k(double K0, double K2, double K3, double K4, double K5, double K6,
                                               double (&q)[4]) {
    q[0] = K0;
    q[1] = fma(TAB(P71), K6, fma(TAB(P61), K5, fma(TAB(P51), K4, fma(TAB(P41), K3, fma(TAB(P31), K2, TAB(P11) * K0)))));
    q[2] = fma(TAB(P72), K6, fma(TAB(P62), K5, fma(TAB(P52), K4, fma(TAB(P42), K3, fma(TAB(P32), K2, TAB(P12) * K0)))));
    q[3] = fma(TAB(P73), K6, fma(TAB(P63), K5, fma(TAB(P53), K4, fma(TAB(P43), K3, fma(TAB(P33), K2, TAB(P13) * K0)))));
}

__device__ __forceinline__ void dense_coeffs_x(double ki, double K0, double K1, double K2, double K3, double K4,
                                               double K5, double h, double (&q)[4]) {
    const double s0 = fma(TAB(PA60), K5, fma(TAB(PA50), K4, fma(TAB(PA40), K3, fma(TAB(PA30), K2, fma(TAB(PA20), K1, TAB(PA10) * K0)))));
    const double s1 = fma(TAB(PA61), K5, fma(TAB(PA51), K4, fma(TAB(PA41), K3, fma(TAB(PA31), K2, fma(TAB(PA21), K1, TAB(PA11) * K0)))));
    const double s2 = fma(TAB(PA62), K5, fma(TAB(PA52), K4, fma(TAB(PA42), K3, fma(TAB(PA32), K2, fma(TAB(PA22), K1, TAB(PA12) * K0)))));
    const double s3 = fma(TAB(PA63), K5, fma(TAB(PA53), K4, fma(TAB(PA43), K3, fma(TAB(PA33), K2, fma(TAB(PA23), K1, TAB(PA13) * K0)))));
    q[0] = fma(h, s0, ki);  // PS[0] = 1
    q[1] = fma(h, s1, TAB(PS1) * ki);
    q[2] = fma(h, s2, TAB(PS2) * ki);
    q[3] = fma(h, s3, TAB(PS3) * ki);
}

// Shared-weight form for evaluating many components at one s: y_i(s) = y_i + h sum_j w_j(s) K_j[i] with
//   momentum:  w_j = sum_c P[j][c] s^(c+1)                     (7 weights, j = 1 has P = 0)
//   position:  x_i(s) = x_i + h (k_i ps(s) + h sum_l v_l(s) K_l[i]),  v_l = sum_c PA[l][c] s^(c+1), ps = sum_c PS[c] s^(c+1)
struct DenseWeights {
    double wk[7];  // wk[1] unused
    double vx[6];
    double ps;
};

__device__ __forceinline__ void dense_weights(double s, DenseWeights& w) {
    const double s2 = s * s;
    w.wk[0] = fma(s2, fma(s, fma(s, TAB(P13), TAB(P12)), TAB(P11)), s);
    w.wk[1] = 0.0;
    w.wk[2] = s2 * fma(s, fma(s, TAB(P33), TAB(P32)), TAB(P31));
    w.wk[3] = s2 * fma(s, fma(s, TAB(P43), TAB(P42)), TAB(P41));
    w.wk[4] = s2 * fma(s, fma(s, TAB(P53), TAB(P52)), TAB(P51));
    w.wk[5] = s2 * fma(s, fma(s, TAB(P63), TAB(P62)), TAB(P61));
    w.wk[6] = s2 * fma(s, fma(s, TAB(P73), TAB(P72)), TAB(P71));
    w.vx[0] = s * fma(s, fma(s, fma(s, TAB(PA13), TAB(PA12)), TAB(PA11)), TAB(PA10));
    w.vx[1] = s * fma(s, fma(s, fma(s, TAB(PA23), TAB(PA22)), TAB(PA21)), TAB(PA20));
    w.vx[2] = s * fma(s, fma(s, fma(s, TAB(PA33), TAB(PA32)), TAB(PA31)), TAB(PA30));
    w.vx[3] = s * fma(s, fma(s, fma(s, TAB(PA43), TAB(PA42)), TAB(PA41)), TAB(PA40));
    w.vx[4] = s * fma(s, fma(s, fma(s, TAB(PA53), TAB(PA52)), TAB(PA51)), TAB(PA50));
    w.vx[5] = s * fma(s, fma(s, fma(s, TAB(PA63), TAB(PA62)), TAB(PA61)), TAB(PA60));
    w.ps = s * fma(s, fma(s, fma(s, TAB(PS3), TAB(PS2)), TAB(PS1)), 1.0);
}

__device__ __forceinline__ double dense_k(const DenseWeights& w, double ki, double K0, double K2, double K3, double K4,
                                          double K5, double K6, double h) {
    const double acc = fma(w.wk[6], K6, fma(w.wk[5], K5, fma(w.wk[4], K4, fma(w.wk[3], K3, fma(w.wk[2], K2, w.wk[0] * K0)))));
    return fma(h, acc, ki);
}

__device__ __forceinline__ double dense_x(const DenseWeights& w, double xi, double ki, double K0, double K1, double K2,
                                          double K3, double K4, double K5, double h) {
    const double acc = fma(w.vx[5], K5, fma(w.vx[4], K4, fma(w.vx[3], K3, fma(w.vx[2], K2, fma(w.vx[1], K1, w.vx[0] * K0)))));
    return fma(h, fma(h, acc, ki * w.ps), xi);
}

__device__ __forceinline__ double dense_eval(const double (&q)[4], double yold, double h, double s) {
    // same term order as numpy: p = cumprod([s,s,s,s]); y = h * dot(Q, p) + y_old
    const double p2 = s * s, p3 = p2 * s, p4 = p3 * s;
    const double acc = fma(q[3], p4, fma(q[2], p3, fma(q[1], p2, q[0] * s)));
    return fma(h, acc, yold);
}

// Root of r(s) - target on s in [0, s_hi] given a sign change between the end points.
// scipy uses brentq(xtol=rtol=4 eps) on the same quartic (ivp.py:52-77); any bracketing method that
// converges to the last bit of s lands inside brentq's own tolerance.  Halley iteration (cubic: typically
// 3 iterations from the secant start) with a bisection safeguard on the maintained bracket.
__device__ __forceinline__ double event_root(const double (&q)[4], double rold, double h, double target,
                                             double s_hi = 1.0) {
    double lo = 0.0, hi = s_hi;
    const double flo = rold - target;
    const double fhi = dense_eval(q, rold, h, s_hi) - target;
    if (flo == 0.0) return 0.0;
    if (fhi == 0.0) return s_hi;
    const bool lo_neg = flo < 0.0;
    double s = s_hi * (flo * fast_rcp5(flo - fhi));  // secant start
    if (!(s > 0.0 && s < s_hi)) s = 0.5 * s_hi;
    const double q1x2 = 2.0 * q[1], q2x3 = 3.0 * q[2], q3x4 = 4.0 * q[3], q2x6 = 6.0 * q[2], q3x12 = 12.0 * q[3];
    for (int it = 0; it < 60; it++) {
        const double fs = dense_eval(q, rold, h, s) - target;
        if (fs == 0.0) return s;
        if ((fs < 0.0) == lo_neg) lo = s; else hi = s;
        const double d1 = h * fma(s, fma(s, fma(s, q3x4, q2x3), q1x2), q[0]);  // f'
        const double d2 = h * fma(s, fma(s, q3x12, q2x6), q1x2);               // f''
        const double den = fma(2.0 * d1, d1, -fs * d2);
        double sn = fma(-2.0 * fs * d1, fast_rcp5(den), s);
        if (!(sn > lo && sn < hi)) sn = 0.5 * (lo + hi);
        if (fabs(sn - s) <= 2.220446049250313e-16 * fmax(fabs(sn), 1e-3) || hi - lo <= 4.4e-16 * hi) return sn;
        s = sn;
    }
    return s;
}

}  // namespace bhg
