// Per-ray FP64 arithmetic of the Schwarzschild null-geodesic path, written for sm_100a.
//
// What it replaces (reference = /root/reference, bldevries/blackhole_geodesic_calculator):
//   curvedpy's calc_trajectory / ray_trace as called at raytracer/RelativisticRenderEngine.py:293-294
//   and raytracer/LimitedRelativisticRenderEngine.py:273-278, i.e. the geodesic RHS of README.md:198-209
//   (metric README.md:162-172, Christoffels README.md:133-135) under scipy's RK45 (README.md:196).
// The step controller, tolerances, initial step and event semantics follow scipy 1.18.1
// (_ivp/rk.py:8-11,14-71,85-176,538-566,715-738; _ivp/common.py:63-134; _ivp/ivp.py:52-158,659-699);
// the arithmetic itself is organised for the FP64 pipe: one reciprocal per RHS, FMA-form stage sums,
// Newton-refined reciprocal / inverse tenth root instead of IEEE div / pow.  Nothing here is translated
// from the reference (it contains no solver code); the CPU restatement lives in oracle/.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

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
    LANE_PENDING_EVENT = -2,  // accepted step crossed an event surface; K, y_old, h kept for the deferred finish
    LANE_EMPTY = -4,
};

// ---------------------------------------------------------------------------------------------
// small FP64 building blocks
// ---------------------------------------------------------------------------------------------

// 1/a to ~1 ulp for normal, finite a (all call sites guarantee that or produce NaN/inf that the step
// controller rejects): MUFU.RCP64H seed (>= 20 bits) + one third-order refinement = 3 DFMA.
__device__ __forceinline__ double fast_rcp(double a) {
    double x0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(a));
    double e = fma(-a, x0, 1.0);
    double e2 = fma(e, e, e);
    double x1 = fma(x0, e2, x0);
    // one more Newton step costs 2 DFMA and makes the result independent of the seed's exact accuracy
    double e3 = fma(-a, x1, 1.0);
    return fma(x1, e3, x1);
}

// a^(-1/10) for a in [1e-12, 1e8]: float seed + two Newton steps on x^-10 = a (quadratic; 1e-6 -> 1e-22)
__device__ __forceinline__ double inv_tenth_root(double a) {
    float af = (float)a;
    double x = (double)exp2f(-0.1f * log2f(af));
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double x2 = x * x;
        double x4 = x2 * x2;
        double x5 = x4 * x;
        double x10 = x5 * x5;
        double r = fma(-a, x10, 11.0);
        x = x * r * 0.1;
    }
    return x;
}

// sin and cos of a moderate argument.  |theta| stays O(pi) on this path; the library routine's
// Payne-Hanek slow path is kept for safety (never taken in practice, costs only a predicate).
__device__ __forceinline__ void sincos_pi(double th, double* s, double* c) {
    sincos(th, s, c);
}

// ---------------------------------------------------------------------------------------------
// RK45 (Dormand-Prince) tableau, scipy/_ivp/rk.py:538-566
// ---------------------------------------------------------------------------------------------
#define BHG_A21 (1.0 / 5)
#define BHG_A31 (3.0 / 40)
#define BHG_A32 (9.0 / 40)
#define BHG_A41 (44.0 / 45)
#define BHG_A42 (-56.0 / 15)
#define BHG_A43 (32.0 / 9)
#define BHG_A51 (19372.0 / 6561)
#define BHG_A52 (-25360.0 / 2187)
#define BHG_A53 (64448.0 / 6561)
#define BHG_A54 (-212.0 / 729)
#define BHG_A61 (9017.0 / 3168)
#define BHG_A62 (-355.0 / 33)
#define BHG_A63 (46732.0 / 5247)
#define BHG_A64 (49.0 / 176)
#define BHG_A65 (-5103.0 / 18656)
#define BHG_B1 (35.0 / 384)
#define BHG_B3 (500.0 / 1113)
#define BHG_B4 (125.0 / 192)
#define BHG_B5 (-2187.0 / 6784)
#define BHG_B6 (11.0 / 84)
#define BHG_E1 (-71.0 / 57600)
#define BHG_E3 (71.0 / 16695)
#define BHG_E4 (-71.0 / 1920)
#define BHG_E5 (17253.0 / 339200)
#define BHG_E6 (-22.0 / 525)
#define BHG_E7 (1.0 / 40)

// ---------------------------------------------------------------------------------------------
// right-hand sides.  State order [k_t, t, k_r, r, k_th, th, k_ph, ph] (parity, NS=8) and
// [k_t, t, k_r, r, k_ph, ph] (orbital plane theta=pi/2, NS=6).
// ---------------------------------------------------------------------------------------------
template <int NS>
struct Rhs;

template <>
struct Rhs<8> {
    static constexpr int IR = 3;
    __device__ __forceinline__ static void eval(const double (&y)[8], double rs, double (&f)[8]) {
        const double kt = y[0], kr = y[2], r = y[3], kth = y[4], th = y[5], kph = y[6];
        double s, c;
        sincos_pi(th, &s, &c);
        const double rm = r - rs;
        // the only two reciprocals of the RHS; independent of each other so they overlap in the pipe
        const double i_rrm = fast_rcp(r * rm);  // 1 / (r (r - rs))
        const double i_s = fast_rcp(s);         // 1 / sin(theta)
        const double i_r = i_rrm * rm;          // 1 / r
        const double A = rs * i_rrm;            // rs / (r (r - rs))
        const double kph2 = kph * kph;
        const double ang = fma(kph2 * s, s, kth * kth);  // k_th^2 + k_ph^2 sin^2
        f[0] = -(A * kr) * kt;
        f[1] = kt;
        // (rs/(2 r rm)) kr^2 - (rs rm/(2 r^3)) kt^2 + rm * ang
        const double half_A = 0.5 * A;
        const double w = (half_A * rm) * (rm * i_r) * i_r;  // rs rm / (2 r^3) = half_A * rm^2 / r^2
        f[2] = fma(half_A * kr, kr, fma(-w * kt, kt, rm * ang));
        f[3] = kr;
        f[4] = fma(kph2 * s, c, -2.0 * (kr * i_r) * kth);
        f[5] = kth;
        f[6] = -2.0 * kph * fma(kr, i_r, kth * (c * i_s));
        f[7] = kph;
    }
};

template <>
struct Rhs<6> {
    static constexpr int IR = 3;
    __device__ __forceinline__ static void eval(const double (&y)[6], double rs, double (&f)[6]) {
        const double kt = y[0], kr = y[2], r = y[3], kph = y[4];
        const double rm = r - rs;
        const double i_rrm = fast_rcp(r * rm);
        const double i_r = i_rrm * rm;
        const double A = rs * i_rrm;
        const double half_A = 0.5 * A;
        const double w = (half_A * rm) * (rm * i_r) * i_r;
        f[0] = -(A * kr) * kt;
        f[1] = kt;
        f[2] = fma(half_A * kr, kr, fma(-w * kt, kt, rm * (kph * kph)));
        f[3] = kr;
        f[4] = -2.0 * kph * (kr * i_r);
        f[5] = kph;
    }
};

// ---------------------------------------------------------------------------------------------
// one RK45 attempt (rk_step + error estimate, scipy/_ivp/rk.py:14-71,105-109,143-146).
// On entry K[0] = f(y).  Fills K[1..6], yn; returns sum_i (err_i / scale_i)^2.
// ---------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ double ynew_component(const double yi, const double k0, const double k2, const double k3,
                                                 const double k4, const double k5, const double h) {
    double acc = BHG_B1 * k0;
    acc = fma(BHG_B3, k2, acc);
    acc = fma(BHG_B4, k3, acc);
    acc = fma(BHG_B5, k4, acc);
    acc = fma(BHG_B6, k5, acc);
    return fma(h, acc, yi);
}

template <int NS>
__device__ __forceinline__ double rk45_attempt(const double (&y)[NS], double (&K)[7][NS], double (&yn)[NS],
                                               const double h, const double rs, const double rtol,
                                               const double atol) {
    double yt[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) yt[i] = fma(h * BHG_A21, K[0][i], y[i]);
    Rhs<NS>::eval(yt, rs, K[1]);
#pragma unroll
    for (int i = 0; i < NS; i++) yt[i] = fma(h, fma(BHG_A32, K[1][i], BHG_A31 * K[0][i]), y[i]);
    Rhs<NS>::eval(yt, rs, K[2]);
#pragma unroll
    for (int i = 0; i < NS; i++)
        yt[i] = fma(h, fma(BHG_A43, K[2][i], fma(BHG_A42, K[1][i], BHG_A41 * K[0][i])), y[i]);
    Rhs<NS>::eval(yt, rs, K[3]);
#pragma unroll
    for (int i = 0; i < NS; i++)
        yt[i] = fma(h, fma(BHG_A54, K[3][i], fma(BHG_A53, K[2][i], fma(BHG_A52, K[1][i], BHG_A51 * K[0][i]))), y[i]);
    Rhs<NS>::eval(yt, rs, K[4]);
#pragma unroll
    for (int i = 0; i < NS; i++)
        yt[i] = fma(h,
                    fma(BHG_A65, K[4][i],
                        fma(BHG_A64, K[3][i], fma(BHG_A63, K[2][i], fma(BHG_A62, K[1][i], BHG_A61 * K[0][i])))),
                    y[i]);
    Rhs<NS>::eval(yt, rs, K[5]);
#pragma unroll
    for (int i = 0; i < NS; i++) yn[i] = ynew_component<NS>(y[i], K[0][i], K[2][i], K[3][i], K[4][i], K[5][i], h);
    Rhs<NS>::eval(yn, rs, K[6]);
    double esum = 0.0;
#pragma unroll
    for (int i = 0; i < NS; i++) {
        double e = BHG_E1 * K[0][i];
        e = fma(BHG_E3, K[2][i], e);
        e = fma(BHG_E4, K[3][i], e);
        e = fma(BHG_E5, K[4][i], e);
        e = fma(BHG_E6, K[5][i], e);
        e = fma(BHG_E7, K[6][i], e);
        const double scale = fma(fmax(fabs(y[i]), fabs(yn[i])), rtol, atol);
        const double q = (e * h) * fast_rcp(scale);
        esum = fma(q, q, esum);
    }
    return esum;
}

// step-size factor 0.9 * err^(-1/5) clipped to [lo, hi], with err^2 = en2 (scipy/_ivp/rk.py:148-163)
__device__ __forceinline__ double step_factor(double en2, double lo, double hi) {
    if (!(en2 < 1e8)) return lo;   // huge, inf or NaN error norm
    if (en2 < 1e-12) return hi;    // includes en2 == 0
    double f = 0.9 * inv_tenth_root(en2);
    return fmin(hi, fmax(lo, f));
}

// 10 * |nextafter(t, +inf) - t|  for t >= 0 (scipy/_ivp/rk.py:119)
__device__ __forceinline__ double min_step_at(double t) {
    double up = __longlong_as_double(__double_as_longlong(t) + 1);
    return 10.0 * (up - t);
}

// ---------------------------------------------------------------------------------------------
// Hairer's initial step (scipy/_ivp/common.py:109-134, order = 4, direction = +1); one RHS evaluation.
// ---------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ double initial_step(const double (&y)[NS], const double (&f0)[NS], double rs, double rtol,
                                               double atol, double interval, double max_step) {
    if (interval == 0.0) return 0.0;
    double iscale[NS];
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int i = 0; i < NS; i++) {
        iscale[i] = 1.0 / fma(fabs(y[i]), rtol, atol);
        const double a = y[i] * iscale[i], b = f0[i] * iscale[i];
        d0 = fma(a, a, d0);
        d1 = fma(b, b, d1);
    }
    d0 = sqrt(d0 / NS);
    d1 = sqrt(d1 / NS);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    h0 = fmin(h0, interval);
    double y1[NS], f1[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) y1[i] = fma(h0, f0[i], y[i]);
    Rhs<NS>::eval(y1, rs, f1);
    double d2 = 0.0;
#pragma unroll
    for (int i = 0; i < NS; i++) {
        const double a = (f1[i] - f0[i]) * iscale[i];
        d2 = fma(a, a, d2);
    }
    d2 = sqrt(d2 / NS) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 0.2);
    return fmin(fmin(100.0 * h0, h1), fmin(interval, max_step));
}

// ---------------------------------------------------------------------------------------------
// dense output (scipy/_ivp/rk.py:554-566,715-738): y(t_old + x h) = y_old + h * sum_c Q_c x^(c+1)
// ---------------------------------------------------------------------------------------------
#define BHG_P11 (-8048581381.0 / 2820520608)
#define BHG_P12 (8663915743.0 / 2820520608)
#define BHG_P13 (-12715105075.0 / 11282082432)
#define BHG_P31 (131558114200.0 / 32700410799)
#define BHG_P32 (-68118460800.0 / 10900136933)
#define BHG_P33 (87487479700.0 / 32700410799)
#define BHG_P41 (-1754552775.0 / 470086768)
#define BHG_P42 (14199869525.0 / 1410260304)
#define BHG_P43 (-10690763975.0 / 1880347072)
#define BHG_P51 (127303824393.0 / 49829197408)
#define BHG_P52 (-318862633887.0 / 49829197408)
#define BHG_P53 (701980252875.0 / 199316789632)
#define BHG_P61 (-282668133.0 / 205662961)
#define BHG_P62 (2019193451.0 / 616988883)
#define BHG_P63 (-1453857185.0 / 822651844)
#define BHG_P71 (40617522.0 / 29380423)
#define BHG_P72 (-110615467.0 / 29380423)
#define BHG_P73 (69997945.0 / 29380423)

__device__ __forceinline__ void dense_coeffs(double k0, double k2, double k3, double k4, double k5, double k6,
                                             double (&q)[4]) {
    q[0] = k0;
    q[1] = fma(BHG_P71, k6, fma(BHG_P61, k5, fma(BHG_P51, k4, fma(BHG_P41, k3, fma(BHG_P31, k2, BHG_P11 * k0)))));
    q[2] = fma(BHG_P72, k6, fma(BHG_P62, k5, fma(BHG_P52, k4, fma(BHG_P42, k3, fma(BHG_P32, k2, BHG_P12 * k0)))));
    q[3] = fma(BHG_P73, k6, fma(BHG_P63, k5, fma(BHG_P53, k4, fma(BHG_P43, k3, fma(BHG_P33, k2, BHG_P13 * k0)))));
}

__device__ __forceinline__ double dense_eval(const double (&q)[4], double yold, double h, double x) {
    // same term order as numpy: p = cumprod([x,x,x,x]); y = h * dot(Q, p) + y_old
    const double p2 = x * x, p3 = p2 * x, p4 = p3 * x;
    const double acc = fma(q[3], p4, fma(q[2], p3, fma(q[1], p2, q[0] * x)));
    return fma(h, acc, yold);
}

// Root of r(x) - target on x in [0,1] given a sign change between the end points.
// scipy uses brentq(xtol=rtol=4 eps) on the same quartic (ivp.py:52-77); any bracketing method that
// converges to the last bit of x lands inside brentq's own tolerance.  Newton with bisection safeguard.
__device__ __forceinline__ double event_root(const double (&q)[4], double rold, double h, double target) {
    double lo = 0.0, hi = 1.0;
    double flo = rold - target;
    double fhi = dense_eval(q, rold, h, 1.0) - target;
    if (flo == 0.0) return 0.0;
    if (fhi == 0.0) return 1.0;
    const bool lo_neg = flo < 0.0;
    double x = flo / (flo - fhi);  // secant start
    if (!(x > 0.0 && x < 1.0)) x = 0.5;
    for (int it = 0; it < 80; it++) {
        const double fx = dense_eval(q, rold, h, x) - target;
        if (fx == 0.0) return x;
        if ((fx < 0.0) == lo_neg) lo = x; else hi = x;
        // derivative of h * (q0 x + q1 x^2 + q2 x^3 + q3 x^4)
        const double dfx = h * fma(4.0 * q[3], x * x * x, fma(3.0 * q[2], x * x, fma(2.0 * q[1], x, q[0])));
        double xn = x - fx / dfx;
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        if (fabs(xn - x) <= 2.220446049250313e-16 * fmax(fabs(xn), 1e-3) || hi - lo <= 4.4e-16 * hi) return xn;
        x = xn;
    }
    return x;
}

}  // namespace bhg
