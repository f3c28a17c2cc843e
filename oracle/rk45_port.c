/* ORACLE (test infrastructure; never linked into, called by, or shipped with the product).
 *
 * PARITY UNPINNED: the reference holds no golden vectors for this path and its arithmetic lives in
 * the absent third-party package `curvedpy` (see oracle/schwarzschild_ref.py header).  This file is a
 * plain-C, scalar, line-by-line restatement of what that path executes per ray:
 *
 *   curvedpy `calc_trajectory` / `SchwarzschildGeodesic.ray_trace`  (call sites
 *     /root/reference/raytracer/RelativisticRenderEngine.py:293-294,
 *     /root/reference/raytracer/LimitedRelativisticRenderEngine.py:273-278)
 *   -> xyz->spherical conversion of position and tangent (RelativisticRenderEngine.py:289-291)
 *   -> null k_t (time_like=False, RelativisticRenderEngine.py:134; r_s = 2M, :95)
 *   -> scipy.integrate.solve_ivp(method="RK45") (README.md:196,211), scipy 1.18.1:
 *        rk.py:8-11     SAFETY, MIN_FACTOR, MAX_FACTOR
 *        rk.py:14-71    rk_step
 *        rk.py:85-103   RungeKutta.__init__ (f0, select_initial_step, error_exponent)
 *        rk.py:111-176  _step_impl (min_step, clamp, t_bound clip, scale, error norm, accept/reject)
 *        rk.py:538-566  RK45 tableau C, A, B, E, P
 *        rk.py:178-180,715-738  dense output  Q = K^T P ; y(t) = y_old + h Q [x,x^2,x^3,x^4]
 *        common.py:63-65   RMS norm ; common.py:68-134 select_initial_step
 *        ivp.py:134-158    find_active_events ; ivp.py:52-77 brentq(xtol=4eps, rtol=4eps)
 *        ivp.py:80-131     handle_events (earliest terminal root) ; ivp.py:659-699 driver loop
 *        base.py:179-210   step() / finished test ; base.py:156-158 nfev counting
 *   -> RHS: the sympy-simplified geodesic equations (README.md:133-135,162-172,198-209), written here
 *      in the same operation order sympy's printer emits (see schwarzschild_ref.rhs_expressions()).
 *
 * It is validated against oracle/schwarzschild_ref.py (real scipy) by tests/test_oracle.py and is
 * used where the pure-Python oracle would take hours (full-frame subsamples).
 *
 * mode 0: 8-state spherical "parity" algorithm.
 * mode 1: 6-state orbital-plane algorithm (restates the product's optional plane mode so that mode
 *         can be checked step-for-step on the CPU; it is NOT something the reference does).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>
#include <pthread.h>
#include <unistd.h>

enum { ESCAPED = 0, CAPTURED = 1, START_INSIDE_HOLE = 2, LAMBDA_EXHAUSTED = 3, STEP_FAILED = 4 };

#define NMAX 8

/* C = [0,1/5,3/10,4/5,8/9,1] is unused: the RHS is autonomous (no explicit lambda dependence). */
static const double A_[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double B_[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double E_[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
static const double P_[7][4] = {
    {1, -8048581381.0 / 2820520608, 8663915743.0 / 2820520608, -12715105075.0 / 11282082432},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799, -68118460800.0 / 10900136933, 87487479700.0 / 32700410799},
    {0, -1754552775.0 / 470086768, 14199869525.0 / 1410260304, -10690763975.0 / 1880347072},
    {0, 127303824393.0 / 49829197408, -318862633887.0 / 49829197408, 701980252875.0 / 199316789632},
    {0, -282668133.0 / 205662961, 2019193451.0 / 616988883, -1453857185.0 / 822651844},
    {0, 40617522.0 / 29380423, -110615467.0 / 29380423, 69997945.0 / 29380423}};

typedef struct {
    int n;       /* 8 (spherical) or 6 (plane) */
    int ir;      /* index of r in the state */
    double rs;
} sys_t;

/* state [k_t, t, k_r, r, k_th, th, k_ph, ph]; expressions as printed by sympy.simplify */
static void rhs8(const double* y, double rs, double* f) {
    double k_t = y[0], k_r = y[2], r = y[3], k_th = y[4], th = y[5], k_ph = y[6];
    double s = sin(th);
    double rm = r - rs;
    f[0] = -k_r * k_t * rs / (r * rm);
    f[1] = k_t;
    f[2] = (k_r * k_r * (r * r) * rs - k_t * k_t * rs * (rm * rm) +
            2 * (r * r * r) * (rm * rm) * (k_ph * k_ph * (s * s) + k_th * k_th)) /
           (2 * (r * r * r) * rm);
    f[3] = k_r;
    f[4] = k_ph * k_ph * sin(2 * th) / 2 - 2 * k_r * k_th / r;
    f[5] = k_th;
    f[6] = -2 * k_ph * (k_r + k_th * r / tan(th)) / r;
    f[7] = k_ph;
}

/* plane mode: theta = pi/2, k_th = 0 -> state [k_t, t, k_r, r, k_ph, ph] */
static void rhs6(const double* y, double rs, double* f) {
    double k_t = y[0], k_r = y[2], r = y[3], k_ph = y[4];
    double rm = r - rs;
    f[0] = -k_r * k_t * rs / (r * rm);
    f[1] = k_t;
    f[2] = (k_r * k_r * (r * r) * rs - k_t * k_t * rs * (rm * rm) + 2 * (r * r * r) * (rm * rm) * (k_ph * k_ph)) /
           (2 * (r * r * r) * rm);
    f[3] = k_r;
    f[4] = -2 * k_ph * k_r / r;
    f[5] = k_ph;
}

/* Conditioning probe (tests only): with a non-zero jitter seed every RHS evaluation is made "backward-stably
 * wrong" by one ulp: each state entry the formulas read is multiplied by 1 + s * 2^-52 before the evaluation and each
 * momentum derivative by another such factor after it, s in {-1, 0, +1} from a per-ray generator.  That is the
 * rounding freedom any other implementation of the same formulas has (operation order, FMA contraction, a 1-ulp
 * reciprocal or sine; cancellation between the large terms of dk_r is included because the INPUTS move).  How far a
 * ray's exit state moves under this jitter is the accuracy to which two implementations of the reference's
 * method can agree on that ray; tests/ use it as the per-ray tolerance floor. */
static uint64_t g_jitter_seed = 0;
static __thread uint64_t g_jitter_state = 0;

void bhg_oracle_set_jitter(uint64_t seed) { g_jitter_seed = seed; }

static inline uint64_t jitter_next(void) { /* splitmix64 */
    uint64_t z = (g_jitter_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static void rhs(const sys_t* S, const double* y, double* f) {
    if (!g_jitter_seed) {
        if (S->n == 8) rhs8(y, S->rs, f); else rhs6(y, S->rs, f);
        return;
    }
    double yj[NMAX];
    uint64_t r = jitter_next();
    for (int i = 0; i < S->n; i++, r >>= 4) yj[i] = y[i] * (1.0 + (double)((int)(r % 3) - 1) * DBL_EPSILON);
    if (S->n == 8) rhs8(yj, S->rs, f); else rhs6(yj, S->rs, f);
    r = jitter_next();
    for (int i = 0; i < S->n; i += 2, r >>= 4) f[i] *= 1.0 + (double)((int)(r % 3) - 1) * DBL_EPSILON;
    for (int i = 1; i < S->n; i += 2) f[i] = y[i - 1];  /* dx/dlambda = k exactly, as in every implementation */
}

static double rms(const double* x, int n) { /* common.py:63-65 */
    double s = 0;
    for (int i = 0; i < n; i++) s += x[i] * x[i];
    return sqrt(s) / sqrt((double)n);
}

/* common.py:68-134 with direction=+1 */
static double select_initial_step(const sys_t* S, double t0, const double* y0, double t_bound, double max_step,
                                  const double* f0, int order, double rtol, double atol, int* nfev) {
    int n = S->n;
    double interval = fabs(t_bound - t0);
    if (interval == 0.0) return 0.0;
    double scale[NMAX], tmp[NMAX], y1[NMAX], f1[NMAX];
    for (int i = 0; i < n; i++) scale[i] = atol + fabs(y0[i]) * rtol;
    for (int i = 0; i < n; i++) tmp[i] = y0[i] / scale[i];
    double d0 = rms(tmp, n);
    for (int i = 0; i < n; i++) tmp[i] = f0[i] / scale[i];
    double d1 = rms(tmp, n);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (interval < h0) h0 = interval;
    for (int i = 0; i < n; i++) y1[i] = y0[i] + h0 * f0[i];
    rhs(S, y1, f1);
    (*nfev)++;
    for (int i = 0; i < n; i++) tmp[i] = (f1[i] - f0[i]) / scale[i];
    double d2 = rms(tmp, n) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / (order + 1));
    double h = 100 * h0;
    if (h1 < h) h = h1;
    if (interval < h) h = interval;
    if (max_step < h) h = max_step;
    return h;
}

/* r-component of the dense output minus target, at absolute time t (rk.py:715-738) */
typedef struct { double t_old, h, r_old, q[4], target; int use_cos; } dense_r_t;

static double dense_r(const dense_r_t* D, double t) {
    double x = (t - D->t_old) / D->h;
    double p1 = x, p2 = p1 * x, p3 = p2 * x, p4 = p3 * x; /* np.cumprod */
    double acc = D->q[0] * p1 + D->q[1] * p2 + D->q[2] * p3 + D->q[3] * p4;
    double v = D->h * acc + D->r_old;
    if (D->use_cos) return cos(v);   /* plane-crossing event g = cos(theta) */
    return v - D->target;
}

/* Brent's method as scipy/optimize/Zeros/brentq.c (xtol, rtol, maxiter=100) */
static double brentq(const dense_r_t* D, double xa, double xb, double xtol, double rtol) {
    double xpre = xa, xcur = xb, xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis;
    double delta, stry, dpre, dblk;
    fpre = dense_r(D, xpre);
    fcur = dense_r(D, xcur);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    for (int i = 0; i < 100; i++) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre; spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre); /* secant */
            } else {
                dpre = (fpre - fcur) / (xpre - xcur); /* inverse quadratic */
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) {
                spre = scur; scur = stry;
            } else {
                spre = sbis; scur = sbis;
            }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = dense_r(D, xcur);
    }
    return xcur;
}

typedef struct {
    double y[NMAX];  /* state at termination (event root, t_bound, or last good state) */
    double lam;
    int status, nfev, n_accept, n_attempt;
    double disk_xy[2]; /* first equatorial-plane crossing inside [disk_in, disk_out], else NaN */
} ray_out_t;

/* disk annulus for the non-terminal plane-crossing event (parity mode only); r_out <= 0 disables */
typedef struct { double r_in, r_out; } disk_t;

static int all_finite(const double* y, int n) {
    for (int i = 0; i < n; i++) if (!isfinite(y[i])) return 0;
    return 1;
}

/* solve_ivp(fun, (0, lambda_max), y0, RK45, events=[hit_blackhole(dir 0), reached_end(dir +1)]) */
static void dense_all(const sys_t* S, double K[7][NMAX], const double* y_old, double h, double x, double* y) {
    double p1 = x, p2 = p1 * x, p3 = p2 * x, p4 = p3 * x;
    for (int i = 0; i < S->n; i++) {
        double q0 = 0, q1 = 0, q2 = 0, q3 = 0;
        for (int j = 0; j < 7; j++) {
            q0 += K[j][i] * P_[j][0]; q1 += K[j][i] * P_[j][1];
            q2 += K[j][i] * P_[j][2]; q3 += K[j][i] * P_[j][3];
        }
        y[i] = h * (q0 * p1 + q1 * p2 + q2 * p3 + q3 * p4) + y_old[i];
    }
}

static void integrate(const sys_t* S, const double* y0, double r_horizon_ev, double r_sphere, double rtol, double atol,
                      double max_step, double t_bound, const disk_t* disk, ray_out_t* out) {
    const int n = S->n, ir = S->ir;
    double t = 0.0, y[NMAX], f[NMAX], K[7][NMAX], y_new[NMAX], ytmp[NMAX];
    int nfev = 0, n_accept = 0, n_attempt = 0;
    memcpy(y, y0, sizeof(double) * n);
    rhs(S, y, f); nfev++;                                             /* rk.py:94 */
    double h_abs = select_initial_step(S, t, y, t_bound, max_step, f, 4, rtol, atol, &nfev); /* rk.py:96-98 */
    const double err_exp = -1.0 / 5;                                  /* rk.py:102 */
    const int have_outer = isfinite(r_sphere);
    double g_h = y[ir] - r_horizon_ev, g_e = have_outer ? y[ir] - r_sphere : -1.0; /* ivp.py:653 */
    const int have_disk = disk && disk->r_out > 0 && n == 8;
    double g_d = have_disk ? cos(y[5]) : 1.0;
    out->disk_xy[0] = out->disk_xy[1] = NAN;
    int disk_hit = 0;
    int status = -2;
    while (status == -2) {
        /* ---- OdeSolver.step (base.py:179-210) ---- */
        if (t == t_bound) { status = LAMBDA_EXHAUSTED; break; }
        /* ---- _step_impl (rk.py:111-176) ---- */
        double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        if (h_abs > max_step) h_abs = max_step;
        else if (h_abs < min_step) h_abs = min_step;
        int accepted = 0, rejected = 0, failed = 0;
        double h = 0, t_new = t;
        while (!accepted) {
            if (h_abs < min_step) { failed = 1; break; }
            h = h_abs;
            t_new = t + h;
            if (t_new - t_bound > 0) t_new = t_bound;
            h = t_new - t;
            h_abs = fabs(h);
            n_attempt++;
            /* rk_step (rk.py:61-71) */
            memcpy(K[0], f, sizeof(double) * n);
            for (int s = 1; s < 6; s++) {
                for (int i = 0; i < n; i++) {
                    double dy = 0;
                    for (int j = 0; j < s; j++) dy += K[j][i] * A_[s][j];
                    ytmp[i] = y[i] + dy * h;
                }
                rhs(S, ytmp, K[s]); nfev++;
            }
            for (int i = 0; i < n; i++) {
                double acc = 0;
                for (int j = 0; j < 6; j++) acc += K[j][i] * B_[j];
                y_new[i] = y[i] + h * acc;
            }
            rhs(S, y_new, K[6]); nfev++;
            double esum = 0;
            for (int i = 0; i < n; i++) {
                double scale = atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol;
                double e = 0;
                for (int j = 0; j < 7; j++) e += K[j][i] * E_[j];
                e = e * h / scale;
                esum += e * e;
            }
            double error_norm = sqrt(esum) / sqrt((double)n);
            if (error_norm < 1) {
                double factor;
                if (error_norm == 0) factor = 10;
                else factor = fmin(10, 0.9 * pow(error_norm, err_exp));
                if (rejected) factor = fmin(1, factor);
                h_abs *= factor;
                accepted = 1;
            } else if (error_norm >= 1) {
                h_abs *= fmax(0.2, 0.9 * pow(error_norm, err_exp));
                rejected = 1;
            } else {
                /* NaN error norm: python's `nan < 1` is False -> reject branch; max(0.2, nan) = 0.2 */
                h_abs *= 0.2;
                rejected = 1;
            }
        }
        if (failed) { status = STEP_FAILED; break; }
        n_accept++;
        double t_old = t;
        double y_old_r = y[ir];
        double y_old[NMAX];
        memcpy(y_old, y, sizeof(double) * n);
        t = t_new;
        memcpy(y, y_new, sizeof(double) * n);
        memcpy(f, K[6], sizeof(double) * n);
        int finished = (t - t_bound >= 0);
        /* ---- events (ivp.py:678-699) ---- */
        double g_h_new = y[ir] - r_horizon_ev, g_e_new = have_outer ? y[ir] - r_sphere : -1.0;
        int act_h = ((g_h <= 0 && g_h_new >= 0) || (g_h >= 0 && g_h_new <= 0)); /* direction 0 */
        int act_e = have_outer && (g_e <= 0 && g_e_new >= 0);                   /* direction +1 */
        double g_d_new = have_disk ? cos(y[5]) : 1.0;
        int act_d = have_disk && !disk_hit && ((g_d <= 0 && g_d_new >= 0) || (g_d >= 0 && g_d_new <= 0));
        double root_d = INFINITY;
        if (act_d) { /* non-terminal: root of cos(theta(t)) on the dense output */
            dense_r_t D;
            D.t_old = t_old; D.h = t - t_old; D.r_old = y_old[5]; D.target = 0; D.use_cos = 1;
            for (int c = 0; c < 4; c++) {
                double q = 0;
                for (int j = 0; j < 7; j++) q += K[j][5] * P_[j][c];
                D.q[c] = q;
            }
            root_d = brentq(&D, t_old, t, 4 * DBL_EPSILON, 4 * DBL_EPSILON);
        }
        if (act_h || act_e) {
            dense_r_t D;
            D.t_old = t_old; D.h = t - t_old; D.r_old = y_old_r; D.use_cos = 0;
            for (int c = 0; c < 4; c++) {
                double q = 0;
                for (int j = 0; j < 7; j++) q += K[j][ir] * P_[j][c];
                D.q[c] = q;
            }
            double root_h = INFINITY, root_e = INFINITY;
            if (act_h) { D.target = r_horizon_ev; root_h = brentq(&D, t_old, t, 4 * DBL_EPSILON, 4 * DBL_EPSILON); }
            if (act_e) { D.target = r_sphere; root_e = brentq(&D, t_old, t, 4 * DBL_EPSILON, 4 * DBL_EPSILON); }
            double root = root_h <= root_e ? root_h : root_e;   /* earliest terminal root (ivp.py:117-126) */
            status = root_h <= root_e ? CAPTURED : ESCAPED;
            if (act_d && root_d <= root) { /* events before the terminal one still count (ivp.py:117-126) */
                double yc[NMAX];
                dense_all(S, K, y_old, t - t_old, (root_d - t_old) / (t - t_old), yc);
                if (yc[3] >= disk->r_in && yc[3] <= disk->r_out) {
                    double st = sin(yc[5]);
                    out->disk_xy[0] = yc[3] * st * cos(yc[7]); out->disk_xy[1] = yc[3] * st * sin(yc[7]);
                    disk_hit = 1;
                }
            }
            /* y = sol(root) */
            double x = (root - t_old) / (t - t_old);
            double p1 = x, p2 = p1 * x, p3 = p2 * x, p4 = p3 * x;
            for (int i = 0; i < n; i++) {
                double q0 = 0, q1 = 0, q2 = 0, q3 = 0;
                for (int j = 0; j < 7; j++) {
                    q0 += K[j][i] * P_[j][0]; q1 += K[j][i] * P_[j][1];
                    q2 += K[j][i] * P_[j][2]; q3 += K[j][i] * P_[j][3];
                }
                y[i] = (t - t_old) * (q0 * p1 + q1 * p2 + q2 * p3 + q3 * p4) + y_old[i];
            }
            t = root;
            break;
        }
        if (act_d) {
            double yc[NMAX];
            dense_all(S, K, y_old, t - t_old, (root_d - t_old) / (t - t_old), yc);
            if (yc[3] >= disk->r_in && yc[3] <= disk->r_out) {
                double st = sin(yc[5]);
                out->disk_xy[0] = yc[3] * st * cos(yc[7]); out->disk_xy[1] = yc[3] * st * sin(yc[7]);
                disk_hit = 1;
            }
        }
        g_h = g_h_new; g_e = g_e_new; g_d = g_d_new;
        if (finished) { status = LAMBDA_EXHAUSTED; break; }
    }
    if (!all_finite(y, n)) status = STEP_FAILED;
    memcpy(out->y, y, sizeof(double) * n);
    out->lam = t; out->status = status; out->nfev = nfev; out->n_accept = n_accept; out->n_attempt = n_attempt;
}

static void trace_parity(const double* X, const double* Kc, double M, double r_sphere, double rtol, double atol,
                         double max_step, double eps, double lambda_max, const disk_t* disk, double* xo, double* ko,
                         ray_out_t* o) {
    double rs = 2 * M;
    double x = X[0], yv = X[1], z = X[2], kx = Kc[0], ky = Kc[1], kz = Kc[2];
    double rho2 = x * x + yv * yv, r2 = rho2 + z * z, r = sqrt(r2), rho = sqrt(rho2);
    o->status = START_INSIDE_HOLE; o->nfev = 0; o->n_accept = 0; o->n_attempt = 0; o->lam = 0;
    o->disk_xy[0] = o->disk_xy[1] = NAN;
    for (int i = 0; i < 3; i++) { xo[i] = NAN; ko[i] = NAN; }
    if (!(r > rs + eps)) return;
    double th = acos(z / r), ph = atan2(yv, x);
    double xk = x * kx + yv * ky;
    double k_r = (xk + z * kz) / r;
    double k_th = (z * xk - rho2 * kz) / (r2 * rho);
    double k_ph = (x * ky - yv * kx) / rho2;
    double s = sin(th);
    double k_t = r * sqrt(k_r * k_r + (r - rs) * r * (k_th * k_th + k_ph * k_ph * s * s)) / (r - rs);
    double y0[8] = {k_t, 0.0, k_r, r, k_th, th, k_ph, ph};
    if (!all_finite(y0, 8)) { o->status = STEP_FAILED; return; } /* scipy base.py:21 refuses such a y0 */
    sys_t S = {8, 3, rs};
    integrate(&S, y0, rs + eps, r_sphere, rtol, atol, max_step, lambda_max, disk, o);
    const double* y = o->y;
    double st = sin(y[5]), ct = cos(y[5]), sp = sin(y[7]), cp = cos(y[7]);
    double R = y[3];
    xo[0] = R * st * cp; xo[1] = R * st * sp; xo[2] = R * ct;
    double vx = y[2] * st * cp + R * ct * cp * y[4] - R * st * sp * y[6];
    double vy = y[2] * st * sp + R * ct * sp * y[4] + R * st * cp * y[6];
    double vz = y[2] * ct - R * st * y[4];
    double nn = sqrt(vx * vx + vy * vy + vz * vz);
    ko[0] = vx / nn; ko[1] = vy / nn; ko[2] = vz / nn;
}

/* plane mode: orthonormal in-plane basis e1 = x/|x|, e2 = (k - (k.e1) e1)/|..|; phi measured from e1 */
static void trace_plane(const double* X, const double* Kc, double M, double r_sphere, double rtol, double atol,
                        double max_step, double eps, double lambda_max, double* xo, double* ko, ray_out_t* o) {
    double rs = 2 * M;
    double r = sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2]);
    o->status = START_INSIDE_HOLE; o->nfev = 0; o->n_accept = 0; o->n_attempt = 0; o->lam = 0;
    o->disk_xy[0] = o->disk_xy[1] = NAN;
    for (int i = 0; i < 3; i++) { xo[i] = NAN; ko[i] = NAN; }
    if (!(r > rs + eps)) return;
    double e1[3] = {X[0] / r, X[1] / r, X[2] / r};
    double k_r = Kc[0] * e1[0] + Kc[1] * e1[1] + Kc[2] * e1[2];
    double w[3] = {Kc[0] - k_r * e1[0], Kc[1] - k_r * e1[1], Kc[2] - k_r * e1[2]};
    double wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double e2[3];
    if (wn > 0) { e2[0] = w[0] / wn; e2[1] = w[1] / wn; e2[2] = w[2] / wn; }
    else { e2[0] = e2[1] = e2[2] = 0; }  /* purely radial ray: phi never changes */
    double k_ph = wn / r;
    double k_t = r * sqrt(k_r * k_r + (r - rs) * r * (k_ph * k_ph)) / (r - rs);
    double y0[6] = {k_t, 0.0, k_r, r, k_ph, 0.0};
    if (!all_finite(y0, 6)) { o->status = STEP_FAILED; return; }
    sys_t S = {6, 3, rs};
    integrate(&S, y0, rs + eps, r_sphere, rtol, atol, max_step, lambda_max, 0, o);
    const double* y = o->y;
    double sp = sin(y[5]), cp = cos(y[5]), R = y[3];
    double a = y[2] * cp - R * sp * y[4], b = y[2] * sp + R * cp * y[4]; /* tangent components along e1, e2 */
    double nn = sqrt(a * a + b * b);
    for (int i = 0; i < 3; i++) {
        xo[i] = R * (cp * e1[i] + sp * e2[i]);
        ko[i] = (a * e1[i] + b * e2[i]) / nn;
    }
}

typedef struct {
    const double *pos, *dir;
    int64_t n;
    double M, r_sphere, rtol, atol, max_step, eps, lambda_max;
    int mode;
    disk_t disk;
    double *disk_xy;
    double *exit_pos, *exit_dir, *lam;
    int32_t *status, *nfev, *n_accept, *n_attempt;
    int64_t next; /* shared work counter, chunks of 64 rays */
} job_t;

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    for (;;) {
        int64_t b = __atomic_fetch_add(&J->next, 64, __ATOMIC_RELAXED);
        if (b >= J->n) break;
        int64_t e = b + 64 < J->n ? b + 64 : J->n;
        for (int64_t i = b; i < e; i++) {
            ray_out_t o;
            g_jitter_state = g_jitter_seed ^ ((uint64_t)i * 0xD1B54A32D192ED03ULL);
            if (J->mode == 0)
                trace_parity(J->pos + 3 * i, J->dir + 3 * i, J->M, J->r_sphere, J->rtol, J->atol, J->max_step, J->eps,
                             J->lambda_max, &J->disk, J->exit_pos + 3 * i, J->exit_dir + 3 * i, &o);
            else
                trace_plane(J->pos + 3 * i, J->dir + 3 * i, J->M, J->r_sphere, J->rtol, J->atol, J->max_step, J->eps,
                            J->lambda_max, J->exit_pos + 3 * i, J->exit_dir + 3 * i, &o);
            J->status[i] = o.status;
            if (J->nfev) J->nfev[i] = o.nfev;
            if (J->n_accept) J->n_accept[i] = o.n_accept;
            if (J->n_attempt) J->n_attempt[i] = o.n_attempt;
            if (J->lam) J->lam[i] = o.lam;
            if (J->disk_xy) { J->disk_xy[2 * i] = o.disk_xy[0]; J->disk_xy[2 * i + 1] = o.disk_xy[1]; }
        }
    }
    return 0;
}

int bhg_oracle_max_threads(void) {
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

/* C entry point (ctypes): AoS [n,3] in/out.  counters may be NULL.  nthreads<=0 -> all online cores.
 * disk_xy (n x 2, may be NULL): first equatorial-plane crossing with disk_r_in <= r <= disk_r_out (parity mode). */
int bhg_oracle_trace(const double* pos, const double* dir, int64_t n, double M, double r_sphere, double rtol,
                     double atol, double max_step, double eps_horizon, double lambda_max, int mode, int nthreads,
                     double* exit_pos, double* exit_dir, int32_t* status, int32_t* nfev, int32_t* n_accept,
                     int32_t* n_attempt, double* lam, double disk_r_in, double disk_r_out, double* disk_xy) {
    job_t J = {pos, dir, n, M, r_sphere, rtol, atol, max_step, eps_horizon, lambda_max, mode,
               {disk_r_in, disk_xy ? disk_r_out : 0.0}, disk_xy,
               exit_pos, exit_dir, lam, status, nfev, n_accept, n_attempt, 0};
    if (nthreads <= 0) nthreads = bhg_oracle_max_threads();
    if (nthreads > 256) nthreads = 256;
    if (nthreads == 1 || n <= 64) { worker(&J); return 0; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads; i++) { if (pthread_create(&th[started], 0, worker, &J) == 0) started++; }
    if (started == 0) worker(&J);
    for (int i = 0; i < started; i++) pthread_join(th[i], 0);
    return 0;
}
