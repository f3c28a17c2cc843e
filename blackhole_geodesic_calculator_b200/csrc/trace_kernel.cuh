// Batched trace kernel: persistent warps, one thread per ray, rays served from a global work queue.
//
// Replaces the reference's per-ray Python loop body (raytracer/RelativisticRenderEngine.py:237 ->
// spacetime_ray_cast :271-313; raytracer/LimitedRelativisticRenderEngine.py:232 -> blackhole_hit :259-335)
// for a whole frame or tile in one launch.
//
// Scheduling.  Adaptive step counts differ between rays (5 .. 100+ attempts inside one frame), so a warp
// does not own a fixed set of 32 rays.  Every lane owns one ray at a time; lanes whose ray has terminated
// park in a *pending* state that keeps the last step's K-stages in registers.  When at least
// `refill_threshold` lanes are idle (or nobody is running) the warp services all idle lanes together:
// event root + dense output + exit conversion + store for the pending ones, then a warp-aggregated
// atomicAdd on the queue head (ballot/popc rank) hands each idle lane its next ray, which is loaded and
// initialised (entry conversion, null k_t, f0, Hairer initial step).  The divergent finish/init code is
// thereby executed once per service instead of once per finishing lane, and the RK45 attempt itself is
// always executed with all running lanes converged.
#pragma once
#include "geodesic_core.cuh"
#include "raygen.cuh"

namespace bhg {

struct TraceArgs {
    const double* in;      // SOA: 6 planes; AOS: pos[n][3]
    const double* in_dir;  // AOS only: dir[n][3]
    double* out;           // SOA: 6 planes; AOS: pos[n][3]
    double* out_dir;       // AOS only
    int32_t* status;
    int32_t* counters;     // nullable: [0,n) attempts, [n,2n) accepted
    const int32_t* order;  // nullable permutation
    // cost binning (cost_key_kernel ...): queue order by predicted cost, used unless *sort_keep != 0 (the bundle was
    // found coherent in memory order and stays as the caller gave it)
    const int32_t* sorted;
    const int* sort_keep;
    const int* idle_budget_dev;  // cost binning decided the bundle is incoherent: a smaller idle budget (more refills) pays
    unsigned long long* queue_head;
    long long n;
    double rs, r_hor, r_sphere, rtol, atol, max_step, lambda_max;
    double atol_over_rtol, inv_rtol2;  // the attempt's error norm works with scale / rtol (geodesic_core.cuh)
    int has_outer;
    int has_max_step;  // max_step is finite (the default is +inf: the clamp is then skipped by a uniform branch)
    int refill_threshold;  // > 0: service when this many lanes are idle
    int idle_budget;       // used when refill_threshold == 0: service when the idle lane-iterations accumulated
                           // since the last service reach this budget ("ski rental": idle until the waste equals
                           // the price of one service), which adapts to coherent and incoherent batches alike
    int tile_width;  // > 0: queue slots enumerate 4 x 8 pixel tiles of a row-major image of this width
    int coords;      // 1: entry / exit states (and polyline, disk points) are given in ISOTROPIC Cartesian coordinates
    // band = slot / (8 tile_width) without a 64-bit division: (slot * tile_magic) >> tile_shift, exact for slot < 2^31
    unsigned long long tile_magic;
    int tile_shift;
    // Prepared rays (pre-pass, prepare_kernel): initial spherical state, f0 and Hairer's first step of every ray in
    // QUEUE order as planes of double2 (coalesced 16-byte loads on refill); NULL = initialise inside the trace kernel
    double2* prep;
    // Long rays first (pre-pass): queue slots whose predicted step count is >= HOT_ESTIMATE (28) are listed in hot_list (in
    // arrival order), flagged in hot_mask (one bit per slot) and served BEFORE the natural queue, which skips them:
    // a 140-attempt photon-ring ray started in the middle of a launch ends 0.1 ms after everybody else.
    // Band progress for the courier (courier_kernel): when set, every finished ray bumps band_done[idx / band_rays]
    // after its exit state is visible device-wide
    int* band_done;
    long long band_rays;
    // Input side of a sharded frame: ray i of the launch is frame ray ((i / map_band) map_stride + map_first) map_band
    // + i % map_band of the `in` arrays (map_band = 0: identity) - the shard reads its bands in place, no compaction
    long long map_band, map_first, map_stride;
    unsigned long long* hot_count;   // [0] = number of listed slots
    int32_t* hot_list;               // [n]
    unsigned int* hot_mask;          // [(n + 31) / 32]
    // DISK variant only: first crossing of the equatorial plane with disk_r_in <= r <= disk_r_out
    double disk_r_in, disk_r_out;
    double* disk_xy;  // [n][2], NaN = no hit
    // POLY variant only: positions sampled on linspace(0, lambda_max, poly_n) up to the termination time
    int poly_n;
    double poly_dt;       // lambda_max / (poly_n - 1)
    double* poly_xyz;     // [n][poly_n][3]; samples beyond poly_count are left untouched
    int32_t* poly_count;  // [n]
};

// memory layout of the ray buffers.  IN_AOS_F32: the same [n][3] arrays stored as float32 (Blender's mathutils
// vectors are float32, RelativisticRenderEngine.py:181-182,223): converted to FP64 on load, integrated in FP64,
// rounded to float32 on store - half the PCIe / HBM bytes.
constexpr int IN_SOA = 0, IN_AOS = 1, IN_AOS_F32 = 2;
// A NaN entry position marks a primary ray that never meets the sphere of influence (written by
// generate_rays_kernel): it is not integrated, keeps its flat direction and gets this status.
constexpr int MISSED_SPHERE = 5;

// queue slot -> ray index.  With the image hint, 32 consecutive slots (one warp's fetch when it starts empty)
// cover a 4 x 8 pixel tile, whose rays have far more similar step counts than 32 pixels of one row.
__device__ __forceinline__ long long slot_to_ray(const TraceArgs& a, long long slot) {
    if (a.sorted && !__ldg(a.sort_keep)) return (long long)__ldg(a.sorted + slot);
    if (a.order) return (long long)__ldg(a.order + slot);
    if (a.tile_width > 0) {
        // tiles 4 pixels wide x 8 tall (bands of 8 image rows); measured against 8 wide x 4 tall: config 2
        // 3.463 vs 3.504 ms, config 3 equal (profiles/r1m_experiments.txt)
        const long long band_sz = 8LL * a.tile_width;
        const long long band = (long long)(((unsigned long long)slot * a.tile_magic) >> a.tile_shift);
        const int t = (int)(slot - band * band_sz);
        const int tile = t >> 5, l = t & 31;
        return band * band_sz + (long long)(l >> 2) * a.tile_width + (tile << 2) + (l & 3);
    }
    return slot;
}

// pending-event encodings (all < LANE_RUNNING)
constexpr int PEND_H = -2;   // horizon event active in the last step
constexpr int PEND_E = -3;   // outer-sphere event active
constexpr int PEND_HE = -5;  // both

// returns false for a ray flagged as missing the sphere (NaN entry position; k holds its flat direction)
template <int IN>
__device__ __forceinline__ bool load_ray(const TraceArgs& a, long long idx, double (&x)[3], double (&k)[3]) {
    if (a.map_band) {
        const long long q = idx / a.map_band;
        idx = (q * a.map_stride + a.map_first) * a.map_band + (idx - q * a.map_band);
    }
    if (IN == IN_AOS_F32) {
        const float* pf = reinterpret_cast<const float*>(a.in);
        const float* df = reinterpret_cast<const float*>(a.in_dir);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            x[c] = (double)__ldg(pf + 3 * idx + c);
            k[c] = (double)__ldg(df + 3 * idx + c);
        }
    } else if (IN == IN_AOS) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            x[c] = __ldg(a.in + 3 * idx + c);
            k[c] = __ldg(a.in_dir + 3 * idx + c);
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            x[c] = __ldg(a.in + c * a.n + idx);
            k[c] = __ldg(a.in + (3 + c) * a.n + idx);
        }
    }
    return x[0] == x[0];  // false iff NaN
}

// ---- isotropic <-> Schwarzschild-coordinate boundary map (bhgeo.h BHG_COORDS_ISOTROPIC) --------------------------
// The reference's older solver generation (`SchwarzschildGeodesic`, LimitedRelativisticRenderEngine.py:90,273) "uses
// the Schwarzschild metric in cartesian coordinates" (README.md:174); README Fig. 5 / Fig. 6 are reproduced to the
// pixel when those are the ISOTROPIC ones, rho with r = rho (1 + r_s / 4 rho)^2 (tests/golden/readme_fig5_fig6.npz).
// The integration itself always runs in the spherical Schwarzschild chart of README.md:162-172; this is the
// point map at the boundary.  Isotropic space is conformally flat, so the angle between a direction and the radial
// unit vector is the metric angle: the radial component of the coordinate tangent scales by dr/drho = 1 - a^2, the
// tangential one by r/rho = (1 + a)^2, a = r_s / (4 rho); directions are re-normalised (affine rescaling).
__device__ __forceinline__ void iso_to_schw(double rs, double (&x)[3], double (&k)[3]) {
    const double rho2 = fma(x[0], x[0], fma(x[1], x[1], x[2] * x[2]));
    const double rho = sqrt(rho2);
    const double a = 0.25 * rs / rho;
    const double c = (-2.0 * a / (1.0 + a)) * fma(k[0], x[0], fma(k[1], x[1], k[2] * x[2])) / rho2;
    double v[3];
#pragma unroll
    for (int i = 0; i < 3; i++) v[i] = fma(c, x[i], k[i]);
    const double inv = rsqrt(fma(v[0], v[0], fma(v[1], v[1], v[2] * v[2])));
    const double s = (1.0 + a) * (1.0 + a);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        k[i] = v[i] * inv;
        x[i] *= s;
    }
}

// isotropic radius of a Schwarzschild radius r >= rs (inside the horizon the chart ends: clamp the root at 0)
__device__ __forceinline__ double iso_radius(double rs, double r) {
    return 0.5 * (r - 0.5 * rs + sqrt(fmax(r * (r - rs), 0.0)));
}

__device__ __forceinline__ void schw_to_iso(double rs, double (&x)[3], double (&k)[3]) {
    const double r2 = fma(x[0], x[0], fma(x[1], x[1], x[2] * x[2]));
    const double r = sqrt(r2);
    const double rho = iso_radius(rs, r);
    const double a = 0.25 * rs / rho;
    const double c = (2.0 * a / (1.0 - a)) * fma(k[0], x[0], fma(k[1], x[1], k[2] * x[2])) / r2;
    double v[3];
#pragma unroll
    for (int i = 0; i < 3; i++) v[i] = fma(c, x[i], k[i]);
    const double inv = rsqrt(fma(v[0], v[0], fma(v[1], v[1], v[2] * v[2])));
    const double s = rho / r;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        k[i] = v[i] * inv;
        x[i] *= s;
    }
}

// outputs are AoS for IN_AOS (exit positions optional: a.out may be NULL), planes for IN_SOA
template <int IN>
__device__ __forceinline__ void store_ray(const TraceArgs& a, long long idx, const double (&x)[3],
                                          const double (&k)[3], int status, int n_attempt, int n_accept) {
    if (IN == IN_AOS_F32) {
        float* pf = reinterpret_cast<float*>(a.out);
        float* df = reinterpret_cast<float*>(a.out_dir);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (pf) pf[3 * idx + c] = (float)x[c];
            df[3 * idx + c] = (float)k[c];
        }
    } else if (IN == IN_AOS) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (a.out) a.out[3 * idx + c] = x[c];
            a.out_dir[3 * idx + c] = k[c];
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            a.out[c * a.n + idx] = x[c];
            a.out[(3 + c) * a.n + idx] = k[c];
        }
    }
    a.status[idx] = status;
    if (a.counters) {
        a.counters[idx] = n_attempt;
        a.counters[a.n + idx] = n_accept;
    }
}

// ---- entry conversion (RelativisticRenderEngine.py:289-291: Conversions().convert_xyz_to_sph) -----------
// returns false when the ray starts at or inside the capture surface
template <int NK>
__device__ __forceinline__ bool init_state(const double (&x0)[3], const double (&k0)[3], double rs, double r_hor,
                                           double (&k)[NK], double (&x)[NK]);

template <>
__device__ __forceinline__ bool init_state<4>(const double (&x0)[3], const double (&k0)[3], double rs, double r_hor,
                                              double (&k)[4], double (&x)[4]) {
    const double rho2 = fma(x0[0], x0[0], x0[1] * x0[1]);
    const double r2 = fma(x0[2], x0[2], rho2);
    const double r = sqrt(r2), rho = sqrt(rho2);
    if (!(r > r_hor)) return false;
    // reciprocals instead of IEEE divisions (fast_rcp5 is bit-identical to 1/a on the B200 self-test sweep;
    // a zero denominator — entry on the polar axis — still yields a non-finite state, i.e. STEP_FAILED)
    const double ir = fast_rcp5(r);
    const double th = acos(x0[2] * ir);
    const double ph = atan2(x0[1], x0[0]);
    const double xk = fma(x0[0], k0[0], x0[1] * k0[1]);
    const double k_r = fma(x0[2], k0[2], xk) * ir;
    const double k_th = fma(x0[2], xk, -rho2 * k0[2]) * fast_rcp5(r2 * rho);
    const double k_ph = fma(x0[0], k0[1], -x0[1] * k0[0]) * fast_rcp5(rho2);
    const double s2 = rho2 * (ir * ir);  // sin^2(theta)
    const double rm = r - rs;
    // null condition g_mn k^m k^n = 0, future-directed root (time_like=False, RelativisticRenderEngine.py:134)
    const double k_t = r * sqrt(fma(k_r, k_r, rm * r * fma(k_ph * k_ph, s2, k_th * k_th))) * fast_rcp5(rm);
    k[0] = k_t; k[1] = k_r; k[2] = k_th; k[3] = k_ph;
    x[0] = 0.0; x[1] = r; x[2] = th; x[3] = ph;
    return true;
}

// orbital-plane frame: e1 = x/|x|, e2 = unit(k - (k.e1) e1); phi measured from e1 inside the plane
__device__ __forceinline__ void plane_frame(const double (&x0)[3], const double (&k0)[3], double& r, double& k_r,
                                            double& wn, double (&e1)[3], double (&e2)[3]) {
    r = sqrt(fma(x0[0], x0[0], fma(x0[1], x0[1], x0[2] * x0[2])));
    const double ir = 1.0 / r;
#pragma unroll
    for (int c = 0; c < 3; c++) e1[c] = x0[c] * ir;
    k_r = fma(k0[0], e1[0], fma(k0[1], e1[1], k0[2] * e1[2]));
    double w[3];
#pragma unroll
    for (int c = 0; c < 3; c++) w[c] = fma(-k_r, e1[c], k0[c]);
    wn = sqrt(fma(w[0], w[0], fma(w[1], w[1], w[2] * w[2])));
    const double iw = wn > 0.0 ? 1.0 / wn : 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) e2[c] = w[c] * iw;
}

template <>
__device__ __forceinline__ bool init_state<3>(const double (&x0)[3], const double (&k0)[3], double rs, double r_hor,
                                              double (&k)[3], double (&x)[3]) {
    double r, k_r, wn, e1[3], e2[3];
    plane_frame(x0, k0, r, k_r, wn, e1, e2);
    if (!(r > r_hor)) return false;
    const double k_ph = wn / r;
    const double rm = r - rs;
    const double k_t = r * sqrt(fma(k_r, k_r, rm * r * (k_ph * k_ph))) / rm;
    k[0] = k_t; k[1] = k_r; k[2] = k_ph;
    x[0] = 0.0; x[1] = r; x[2] = 0.0;
    return true;
}

// ---- exit conversion: spherical state -> Cartesian position + unit direction ----------------------------
template <int NK>
__device__ __forceinline__ void exit_state(const double (&k)[NK], const double (&x)[NK], const double (&x0)[3],
                                           const double (&k0)[3], double (&xo)[3], double (&ko)[3]);

template <>
__device__ __forceinline__ void exit_state<4>(const double (&k)[4], const double (&x)[4], const double (&)[3],
                                              const double (&)[3], double (&xo)[3], double (&ko)[3]) {
    double st, ct, sp, cp;
    sincos_tab(x[2], &st, &ct);
    sincos_tab(x[3], &sp, &cp);
    const double R = x[1], k_r = k[1], k_th = k[2], k_ph = k[3];
    xo[0] = R * st * cp;
    xo[1] = R * st * sp;
    xo[2] = R * ct;
    const double a = fma(k_r, st, R * ct * k_th);  // d(rho)/dlambda
    const double b = R * st * k_ph;                // rho * dphi/dlambda
    double v[3];
    v[0] = fma(a, cp, -b * sp);
    v[1] = fma(a, sp, b * cp);
    v[2] = fma(k_r, ct, -R * st * k_th);
    const double inv = rsqrt(fma(v[0], v[0], fma(v[1], v[1], v[2] * v[2])));
#pragma unroll
    for (int c = 0; c < 3; c++) ko[c] = v[c] * inv;
}

template <>
__device__ __forceinline__ void exit_state<3>(const double (&k)[3], const double (&x)[3], const double (&x0)[3],
                                              const double (&k0)[3], double (&xo)[3], double (&ko)[3]) {
    double r0, kr0, wn, e1[3], e2[3];
    plane_frame(x0, k0, r0, kr0, wn, e1, e2);
    double sp, cp;
    sincos_tab(x[2], &sp, &cp);
    const double R = x[1];
    const double a = fma(k[1], cp, -R * sp * k[2]);
    const double b = fma(k[1], sp, R * cp * k[2]);
    const double inv = rsqrt(fma(a, a, b * b));
#pragma unroll
    for (int c = 0; c < 3; c++) {
        xo[c] = R * fma(cp, e1[c], sp * e2[c]);
        ko[c] = fma(a, e1[c], b * e2[c]) * inv;
    }
}

// every component finite?  0 * v is NaN exactly for v = +-inf / NaN: two short FMA chains and one compare instead of
// 2 NK exponent tests (the service path is latency-bound: profiles/r2a_instruction_budget_by_source.txt)
template <int NK>
__device__ __forceinline__ bool all_finite(const double (&k)[NK], const double (&x)[NK]) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int i = 0; i < NK; i++) {
        a = fma(k[i], 0.0, a);
        b = fma(x[i], 0.0, b);
    }
    return (a + b) == 0.0;
}

// Equatorial-plane crossing inside the step [0, s_max] of the dense output (non-terminal event):
// the continuous form of the reference's checkHitDisk polyline scan (LimitedRelativisticRenderEngine.py:413-438:
// z sign change -> crossing point -> R_in <= R <= R_out).  z = r cos(theta) changes sign where theta crosses
// pi/2 + m pi; the crossing is located on the quartic interpolant of theta exactly like the terminal events.
// Returns true (and writes the in-plane hit point) for a crossing inside the annulus.
__device__ __forceinline__ bool disk_crossing(const TraceArgs& a, long long idx, const double (&k)[4],
                                              const double (&x)[4], const double (&K)[7][4], double h, double s_max) {
    const double half_pi = 1.57079632679489661923, inv_pi = 0.31830988618379067154, pi = 3.14159265358979323846;
    double q[4];
    dense_coeffs_x(k[2], K[0][2], K[1][2], K[2][2], K[3][2], K[4][2], K[5][2], h, q);
    const double th_end = dense_eval(q, x[2], h, s_max);
    const double m0 = floor((x[2] - half_pi) * inv_pi), m1 = floor((th_end - half_pi) * inv_pi);
    if (m0 == m1) return false;
    const double m = m1 > m0 ? m0 + 1.0 : m0;  // first cell boundary in the direction of motion
    const double target = fma(m, pi, half_pi);
    const double s = event_root(q, x[2], h, target, s_max);
    double qr[4], qp[4];
    dense_coeffs_x(k[1], K[0][1], K[1][1], K[2][1], K[3][1], K[4][1], K[5][1], h, qr);
    const double r = dense_eval(qr, x[1], h, s);
    if (!(r >= a.disk_r_in && r <= a.disk_r_out)) return false;
    dense_coeffs_x(k[3], K[0][3], K[1][3], K[2][3], K[3][3], K[4][3], K[5][3], h, qp);
    const double ph = dense_eval(qp, x[3], h, s);
    double sp, cp;
    sincos_tab(ph, &sp, &cp);
    const double st = (((long long)m) & 1) ? -1.0 : 1.0;  // sin(pi/2 + m pi)
    const double ro = a.coords ? iso_radius(a.rs, r) : r;
    a.disk_xy[2 * idx] = ro * st * cp;
    a.disk_xy[2 * idx + 1] = ro * st * sp;
    return true;
}

// One 512-thread block per SM (16 warps, 128 registers/thread).  Measured on B200 with the same 16 warps/SM:
// 32x16: 3.514, 64x8: 3.517, 128x4: 3.517, 256x2: 3.493, 512x1: 3.458 ms/frame (profiles/r1m_experiments.txt).
// Polyline samples (second half): the reference asks curvedpy for nr_points_curve samples on
// linspace(0, curve_end, N) and receives those up to the termination time (RelativisticRenderEngine.py:293-294;
// solve_ivp t_eval semantics, scipy/_ivp/ivp.py:711-728).  Emits every pending sample time t_j <= t_end inside the
// step that started at t_old with state (k, x): dense output of r, theta, phi -> Cartesian position.
__device__ __forceinline__ int emit_polyline(const TraceArgs& a, long long idx, int pj, const double (&k)[4],
                                             const double (&x)[4], const double (&K)[7][4], double t_old, double h,
                                             double t_end) {
    while (pj < a.poly_n) {
        const double tj = (pj == a.poly_n - 1) ? a.lambda_max : pj * a.poly_dt;  // numpy.linspace end point is exact
        if (!(tj <= t_end)) break;
        double r = x[1], th = x[2], ph = x[3];
        if (tj > t_old) {
            const double s = (tj - t_old) / h;
            DenseWeights w;
            dense_weights(s, w);
            r = dense_x(w, x[1], k[1], K[0][1], K[1][1], K[2][1], K[3][1], K[4][1], K[5][1], h);
            th = dense_x(w, x[2], k[2], K[0][2], K[1][2], K[2][2], K[3][2], K[4][2], K[5][2], h);
            ph = dense_x(w, x[3], k[3], K[0][3], K[1][3], K[2][3], K[3][3], K[4][3], K[5][3], h);
        }
        double st, ct, sp, cp;
        sincos_tab(th, &st, &ct);
        sincos_tab(ph, &sp, &cp);
        double* o = a.poly_xyz + ((long long)idx * a.poly_n + pj) * 3;
        if (a.coords) r = iso_radius(a.rs, r);
        o[0] = r * st * cp;
        o[1] = r * st * sp;
        o[2] = r * ct;
        pj++;
    }
    return pj;
}

// ---- predicted step count of a ray from its entry state ---------------------------------------------------------------
// Two quantities govern the adaptive step count (fitted to configs 2 and 5 with the oracle,
// profiles/r2g_cost_predictor.txt): s = b^2 / 27 M^2 - 1, the distance of the impact parameter b = L / E from the
// critical one (the photon-sphere winding: +2.2 attempts per halving of |s| for escaping rays, +1.2 for captured ones),
// and |n_z|, the polar component of the unit normal of the orbital plane: the plane passes the coordinate pole at
// sin(theta_min) = |n_z|, where the reference's spherical formulation takes small steps (+2.3 attempts per halving
// for a fly-by, +7 for a ray that winds).  b alone predicts nothing for the 96 % of rays far from critical.
// Float arithmetic: the estimate only orders the queue, it never touches a result.
__device__ __forceinline__ float estimate_attempts(const double (&x)[3], const double (&k)[3], double rs) {
    const float x0 = (float)x[0], x1 = (float)x[1], x2 = (float)x[2];
    const float k0 = (float)k[0], k1 = (float)k[1], k2 = (float)k[2];
    const float lx = x1 * k2 - x2 * k1, ly = x2 * k0 - x0 * k2, lz = x0 * k1 - x1 * k0;
    const float l2 = lx * lx + ly * ly + lz * lz;
    const float r2 = x0 * x0 + x1 * x1 + x2 * x2;
    const float xk = x0 * k0 + x1 * k1 + x2 * k2;
    const float r = sqrtf(r2);
    const float b2 = l2 * r2 / (xk * xk + (1.0f - (float)rs / r) * l2);   // (L / E)^2
    const float sgn = b2 * (float)(1.0 / (6.75 * rs * rs)) - 1.0f;        // 27 M^2 = 6.75 rs^2
    const float u = fabsf(sgn);
    const float nz = fabsf(lz) * rsqrtf(l2);
    float base = 12.0f;
    if (u < 2.0f) {
        const float lg = 1.0f - __log2f(fmaxf(u, 1e-7f));   // log2(2 / u) >= 0
        base = sgn > 0.0f ? 19.8f + 2.2f * lg : 33.3f + 1.2f * lg;
    }
    float est = base;
    if (nz < 0.5f) est += (2.26f + 0.158f * (base - 12.0f)) * (-1.0f - __log2f(fmaxf(nz, 1e-7f)));
    return est < 1e6f ? est : 12.0f;   // NaN / inf from a degenerate entry state
}
constexpr float HOT_ESTIMATE = 28.0f;   // 1.7 % of a config-2 frame, 249 of its 253 rays above 50 attempts (profiles/r2g_cost_predictor.txt)

// ---- prepared rays -------------------------------------------------------------------------------------------------
// Record of one ray after the pre-pass, NK = 4: {k_t, k_r, k_th, k_ph, r, th, ph, K0[4], h0} = 12 doubles = 6 double2
// planes; NK = 3: {k_t, k_r, k_ph, r, K0[3], h0} = 8 doubles = 4 planes (t = 0, plane phi = 0).  h0 < 0 encodes the rays
// that are not integrated: -1 START_INSIDE_HOLE, -2 STEP_FAILED (singular entry), -3 MISSED_SPHERE (k holds the flat
// direction).
template <int NK>
struct Prep {
    static constexpr int PL = (NK == 4) ? 6 : 4;
    __device__ __forceinline__ static void pack(const double (&k)[NK], const double (&x)[NK], const double (&K0)[NK],
                                                double h0, double (&v)[2 * PL]) {
        int j = 0;
#pragma unroll
        for (int i = 0; i < NK; i++) v[j++] = k[i];
#pragma unroll
        for (int i = 1; i < (NK == 4 ? 4 : 2); i++) v[j++] = x[i];
#pragma unroll
        for (int i = 0; i < NK; i++) v[j++] = K0[i];
        v[j] = h0;
    }
    __device__ __forceinline__ static double unpack(const double (&v)[2 * PL], double (&k)[NK], double (&x)[NK],
                                                    double (&K0)[NK]) {
        int j = 0;
#pragma unroll
        for (int i = 0; i < NK; i++) k[i] = v[j++];
        x[0] = 0.0;
#pragma unroll
        for (int i = 1; i < NK; i++) x[i] = (NK == 4 || i < 2) ? v[j++] : 0.0;
#pragma unroll
        for (int i = 0; i < NK; i++) K0[i] = v[j++];
        return v[j];
    }
};

constexpr int IN_CAMERA = 3;  // prepare_kernel only: rays come from the camera description, not from memory

// Pre-pass: entry conversion (xyz -> spherical position and tangent, null k_t: RelativisticRenderEngine.py:289-291,134),
// f0 and Hairer's initial step (scipy/_ivp/rk.py:94-103, common.py:68-134) for every ray, written in queue order.
// In the trace kernel this work runs once per warp-service on a latency-bound dependency chain (~1700 instructions,
// 12 % of all warp time); here it is throughput-bound streaming work for 2048 resident threads per SM.
template <int NK, int IN>
__global__ void __launch_bounds__(256, 4) prepare_kernel(const TraceArgs a, const Camera cam, double* __restrict__ ray_pos,
                                                      double* __restrict__ ray_dir) {
    constexpr int PL = Prep<NK>::PL;
    for (long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x; slot < a.n;
         slot += (long long)gridDim.x * blockDim.x) {
        const long long idx = slot_to_ray(a, slot);
        double x0[3], k0[3];
        bool enters;
        if (IN == IN_CAMERA) {
            enters = camera_ray(cam, idx, x0, k0);
            if (ray_pos) {  // plane mode rebuilds its orbital frame from the flat entry state at the exit
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    ray_pos[3 * idx + c] = enters ? x0[c] : __longlong_as_double(0x7ff8000000000000LL);
                    ray_dir[3 * idx + c] = k0[c];
                }
            }
        } else {
            enters = load_ray<(IN == IN_CAMERA ? IN_AOS : IN)>(a, idx, x0, k0);
        }
        double k[NK], x[NK], K0[NK], h0;
#pragma unroll
        for (int i = 0; i < NK; i++) k[i] = x[i] = K0[i] = 0.0;
        if (!enters) {
            h0 = -3.0;
#pragma unroll
            for (int c = 0; c < 3; c++) k[c] = k0[c];
        } else {
            if (a.coords) iso_to_schw(a.rs, x0, k0);
            const float est = a.hot_list ? estimate_attempts(x0, k0, a.rs) : 0.0f;
            if (!init_state<NK>(x0, k0, a.rs, a.r_hor, k, x)) {
                h0 = -1.0;
            } else if (!all_finite<NK>(k, x)) {
                h0 = -2.0;
            } else {
                if (est >= HOT_ESTIMATE) {
                    const unsigned long long j = atomicAdd(a.hot_count, 1ULL);
                    a.hot_list[j] = (int32_t)slot;
                    atomicOr(a.hot_mask + (slot >> 5), 1u << (slot & 31));
                }
                Rhs<NK>::eval(k, x, a.rs, K0);
                h0 = initial_step<NK>(k, x, K0, a.rs, a.rtol, a.atol, a.lambda_max, a.max_step);
                if (h0 < 0.0) h0 = 0.0;  // cannot happen (steps are non-negative); keeps the status encoding unambiguous
            }
        }
        double v[2 * PL];
        Prep<NK>::pack(k, x, K0, h0, v);
#pragma unroll
        for (int pl = 0; pl < PL; pl++) a.prep[(long long)pl * a.n + slot] = make_double2(v[2 * pl], v[2 * pl + 1]);
    }
}

// ---- cost binning for bundles without image order -------------------------------------------------------------------
// Rays are sorted into COST_BINS classes of estimate_attempts(), costliest first: a warp's 32 lanes then carry similar
// step counts, so the queue's lanes idle less (config 5 in random planes: -5 %).  A bundle whose memory order is
// already coherent (neighbouring rays of an image: their step counts agree far better than any estimate) must be left
// alone (+16 % when sorted): a first pass classifies every 16th group of 32 consecutive rays and, if most groups span
// at most two neighbouring classes, sets the keep flag; the full key / scatter passes then return at once and the
// trace kernel serves the caller's order.  All decisions are taken on the device: no host synchronisation.
// hist layout (ints): [0, COST_BINS) counts, [COST_BINS, 2 COST_BINS) cursors, [2 COST_BINS] coherent groups of the
// sample, [2 COST_BINS + 1] sampled groups, [2 COST_BINS + 2] keep flag, [2 COST_BINS + 3] idle budget.
constexpr int COST_BINS = 32;
constexpr float COST_BIN_WIDTH = 4.0f;
constexpr int COST_SAMPLE_STRIDE = 16;

template <int IN>
__device__ __forceinline__ int cost_key(const TraceArgs& a, long long idx) {
    double x[3], k[3];
    if (!load_ray<IN>(a, idx, x, k)) return 0;
    if (a.coords) iso_to_schw(a.rs, x, k);
    return min(COST_BINS - 1, max(0, (int)(estimate_attempts(x, k, a.rs) * (1.0f / COST_BIN_WIDTH)) - 2));
}

// pass 0: coherence of a sample of the bundle (one warp per sampled group)
template <int IN>
__global__ void __launch_bounds__(256) cost_sample_kernel(const TraceArgs a, int* __restrict__ hist) {
    const int lane = threadIdx.x & 31;
    const long long groups = (a.n + 31) / 32;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long g = warp * COST_SAMPLE_STRIDE; g < groups; g += nwarps * COST_SAMPLE_STRIDE) {
        const long long i = g * 32 + lane;
        const int key = i < a.n ? cost_key<IN>(a, a.order ? (long long)__ldg(a.order + i) : i) : -1;
        const int hi = __reduce_max_sync(0xffffffffu, key);
        const int lo = __reduce_min_sync(0xffffffffu, key < 0 ? COST_BINS : key);
        if (lane == 0) {
            atomicAdd(&hist[2 * COST_BINS + 1], 1);
            if (hi - lo <= 1) atomicAdd(&hist[2 * COST_BINS], 1);
        }
    }
}

// [2 COST_BINS + 3] = idle budget of the refill policy: services are cheap since the pre-pass, so a bundle whose lanes
// finish at very different times refills earlier (32 lane-iterations: -4 % on config 5), a coherent one keeps the
// budget tuned for tiles (profiles/r2m_budget_probe.txt)
__global__ void cost_decide_kernel(int* __restrict__ hist, int force, int budget_coherent, int budget_incoherent) {
    if (threadIdx.x == 0) {
        const int keep = (!force && 2 * hist[2 * COST_BINS] > hist[2 * COST_BINS + 1]) ? 1 : 0;
        hist[2 * COST_BINS + 2] = keep;
        hist[2 * COST_BINS + 3] = keep ? budget_coherent : budget_incoherent;
    }
}

// pass 1: class of every ray + histogram
template <int IN>
__global__ void __launch_bounds__(256) cost_key_kernel(const TraceArgs a, unsigned char* __restrict__ keys,
                                                       int* __restrict__ hist) {
    if (hist[2 * COST_BINS + 2]) return;
    __shared__ int s_hist[COST_BINS];
    if (threadIdx.x < COST_BINS) s_hist[threadIdx.x] = 0;
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        const int key = cost_key<IN>(a, a.order ? (long long)__ldg(a.order + i) : i);
        keys[i] = (unsigned char)key;
        atomicAdd(&s_hist[key], 1);
    }
    __syncthreads();
    if (threadIdx.x < COST_BINS && s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
}

// cursor[key] = first queue slot of bin `key`, costliest bin first
__global__ void cost_offsets_kernel(int* __restrict__ hist) {
    if (threadIdx.x == 0 && !hist[2 * COST_BINS + 2]) {
        int acc = 0;
        for (int key = COST_BINS - 1; key >= 0; key--) {
            hist[COST_BINS + key] = acc;
            acc += hist[key];
        }
    }
}

// pass 2: queue order
__global__ void __launch_bounds__(256) cost_scatter_kernel(long long n, const int32_t* __restrict__ user_order,
                                                          const unsigned char* __restrict__ keys, int* __restrict__ hist,
                                                          int32_t* __restrict__ sorted) {
    if (hist[2 * COST_BINS + 2]) return;
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    int* cursor = hist + COST_BINS;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {
        const long long i = i0 + lane;
        const bool valid = i < n;
        const int key = valid ? (int)keys[i] : -1;
        // lanes of the same bin share one atomic (consecutive slots inside the bin keep the warp's ray order)
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (valid && lane == leader) base = atomicAdd(&cursor[key], __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (valid) sorted[base + __popc(peers & ((1u << lane) - 1u))] = user_order ? __ldg(user_order + i) : (int32_t)i;
    }
}

#ifndef BHG_MIN_BLOCKS
#define BHG_MIN_BLOCKS 1
#endif
#ifndef BHG_BLOCK
#define BHG_BLOCK 512
#endif

// STAGE: the launcher selects it when the output buffers live in another GPU's memory (it costs 0.6 % on local ones)
template <int NK, int IN, bool DISK = false, bool POLY = false, bool PREP = false, bool STAGE = false>
__global__ void __launch_bounds__(BHG_BLOCK, BHG_MIN_BLOCKS) trace_kernel(const TraceArgs a) {
    static_assert(!(DISK || POLY) || NK == 4, "disk event and polyline are defined for the spherical (parity) state");
    constexpr int IR = 1;  // index of r in x
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;

    // (sin, cos)(n pi / 512) for the attempt's sincos_lut (parity mode; the plane system has no angle in its RHS)
    __shared__ double2 s_sincos[NK == 4 ? SINCOS_LUT_N : 1];
    if (NK == 4) {
        const double2* g = reinterpret_cast<const double2*>(&g_sincos_lut[0][0]);
        for (int i = threadIdx.x; i < SINCOS_LUT_N; i += blockDim.x) s_sincos[i] = __ldg(g + i);
        __syncthreads();
    }
    const double2* lut = (NK == 4) ? s_sincos : nullptr;
    // Exit states of the float64 AoS layout leave through a per-warp staging tile: the lanes that finish in one service
    // put their 6 doubles there and the whole warp writes them out element-wise, so rays that are neighbours in memory
    // (4 per tile row, 32 without the tile mapping) become runs of full 32-byte sectors instead of 8-byte pieces at a
    // 24-byte stride.  Local HBM does not care; a frame owner's memory behind NVLink does (distributed.py "stores").
    constexpr bool STAGED = STAGE && (IN == IN_AOS);
    __shared__ double s_stage[STAGED ? BHG_BLOCK / 32 : 1][STAGED ? 6 : 1][32];
    double (*stage)[32] = s_stage[STAGED ? (threadIdx.x >> 5) : 0];

    double k[NK], x[NK], K[7][NK], kn[NK], xn[NK];
    double t = 0.0, h_abs = 0.0;
    long long idx = -1;
    int state = LANE_EMPTY;
    int n_attempt = 0, n_accept = 0;
    bool rejected = false;
    bool disk_hit = false;
    int pj = 0;  // POLY: next polyline sample index
    bool exhausted = false;  // warp-uniform: queue has no more rays
    const int T = a.refill_threshold;
    const int B = a.idle_budget_dev ? __ldg(a.idle_budget_dev) : a.idle_budget;
    int idle_acc = 0;  // warp-uniform
    const double t_bound = a.lambda_max;
    // queue = [listed long rays][natural slots, listed ones skipped]; a list longer than n / 8 is no tail problem but
    // the bulk of the work (a near-critical bundle): it is then ignored
    long long n_hot = 0;
    if (PREP && a.hot_list) {
        n_hot = (long long)*a.hot_count;
        if (n_hot > (a.n >> 3)) n_hot = -1;
    }
    const bool use_hot = n_hot > 0;
    const long long q_len = a.n + (use_hot ? n_hot : 0);

#pragma unroll
    for (int i = 0; i < NK; i++) {
        k[i] = x[i] = kn[i] = xn[i] = 0.0;
#pragma unroll
        for (int s = 0; s < 7; s++) K[s][i] = 0.0;
    }

    while (true) {
        const unsigned running = __ballot_sync(FULL, state == LANE_RUNNING);
        const unsigned empty = __ballot_sync(FULL, state == LANE_EMPTY);
        const unsigned pending = ~(running | empty);
        bool service;
        if (exhausted) {
            if (running == 0u && pending == 0u) break;
            service = (running == 0u);  // drain: finish everybody together at the very end
        } else {
            const int n_idle = __popc(~running);
            idle_acc += n_idle;
            service = (running == 0u) || (T > 0 ? n_idle >= T : idle_acc >= B);
        }

        if (service) {
            idle_acc = 0;
            // ---------------- finish pending lanes ----------------
            const bool was_finishing = state != LANE_RUNNING && state != LANE_EMPTY;
            const unsigned fin = STAGED ? __ballot_sync(FULL, was_finishing) : 0u;
            const long long fin_idx = idx;
            if (state != LANE_RUNNING && state != LANE_EMPTY) {
                int final_status = state;
                if (state < LANE_RUNNING) {
                    // the last accepted step [t, t + h_abs] crossed an event surface (ivp.py:678-697)
                    const double h = h_abs;
                    double q[4];
                    dense_coeffs_x(k[IR], K[0][IR], K[1][IR], K[2][IR], K[3][IR], K[4][IR], K[5][IR], h, q);
                    // one root search per lane: the horizon for PEND_H / PEND_HE, the sphere for PEND_E; only a step
                    // that crossed BOTH surfaces (never seen outside tests) searches the second one as well
                    const bool first_is_h = (state != PEND_E);
                    double s = event_root(q, x[IR], h, first_is_h ? a.r_hor : a.r_sphere);
                    final_status = first_is_h ? CAPTURED : ESCAPED;
                    if (state == PEND_HE) {
                        const double s_e = event_root(q, x[IR], h, a.r_sphere);
                        if (s_e < s) {  // earliest terminal event wins, the horizon on a tie (ivp.py:117-126)
                            s = s_e;
                            final_status = ESCAPED;
                        }
                    }
                    if constexpr (DISK) {  // crossings before the terminal root still count
                        if (!disk_hit) disk_hit = disk_crossing(a, idx, k, x, K, h, s);
                    }
                    if constexpr (POLY) pj = emit_polyline(a, idx, pj, k, x, K, t, h, fma(s, h, t));
                    // dense output at the event (rk.py:715-738); k_t and t are not needed by the exit conversion
                    DenseWeights w;
                    dense_weights(s, w);
#pragma unroll
                    for (int i = 1; i < NK; i++) {
                        const double xi = dense_x(w, x[i], k[i], K[0][i], K[1][i], K[2][i], K[3][i], K[4][i], K[5][i], h);
                        k[i] = dense_k(w, k[i], K[0][i], K[2][i], K[3][i], K[4][i], K[5][i], K[6][i], h);
                        x[i] = xi;
                    }
                    t = fma(s, h, t);
                }
                double x0[3] = {0, 0, 0}, k0[3] = {0, 0, 0}, xo[3], ko[3];
                if (NK == 3 || (!PREP && final_status == MISSED_SPHERE)) {
                    const bool enters = load_ray<IN>(a, idx, x0, k0);
                    if (a.coords && enters) iso_to_schw(a.rs, x0, k0);
                }
                if (PREP && final_status == MISSED_SPHERE) {  // the record carries the flat direction
#pragma unroll
                    for (int c = 0; c < 3; c++) k0[c] = k[c];
                }
                if (final_status == START_INSIDE_HOLE) {
#pragma unroll
                    for (int c = 0; c < 3; c++) xo[c] = ko[c] = __longlong_as_double(0x7ff8000000000000LL);
                } else if (final_status == MISSED_SPHERE) {
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        xo[c] = __longlong_as_double(0x7ff8000000000000LL);
                        ko[c] = k0[c];  // the ray continues in flat space with its original direction
                    }
                } else {
                    if (!all_finite<NK>(k, x)) final_status = STEP_FAILED;
                    exit_state<NK>(k, x, x0, k0, xo, ko);
                    if (a.coords) schw_to_iso(a.rs, xo, ko);
                }
                if constexpr (STAGED) {
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        stage[c][lane] = xo[c];
                        stage[3 + c][lane] = ko[c];
                    }
                    a.status[idx] = final_status;
                    if (a.counters) {
                        a.counters[idx] = n_attempt;
                        a.counters[a.n + idx] = n_accept;
                    }
                } else {
                    store_ray<IN>(a, idx, xo, ko, final_status, n_attempt, n_accept);
                }
                if constexpr (POLY) a.poly_count[idx] = pj;
                if constexpr (DISK) {
                    if (!disk_hit) a.disk_xy[2 * idx] = a.disk_xy[2 * idx + 1] = __longlong_as_double(0x7ff8000000000000LL);
                }
                state = LANE_EMPTY;
            }
            if constexpr (STAGED) {
                if (fin) {
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const int e = 32 * j + lane;        // element e of the tile: ray slot e / 3, component e % 3
                        const int sl = e / 3;
                        const int c = e - 3 * sl;
                        const long long i = __shfl_sync(FULL, fin_idx, sl);
                        if ((fin >> sl) & 1u) {
                            if (a.out) a.out[3 * i + c] = stage[c][sl];
                            a.out_dir[3 * i + c] = stage[3 + c][sl];
                        }
                    }
                    __syncwarp();
                }
            }
            if (a.band_done) {  // warp-uniform
                const unsigned done = __ballot_sync(FULL, fin_idx >= 0 && was_finishing);
                if (done) {
                    __threadfence();  // this warp's exit states before the counters
                    const long long band = was_finishing ? fin_idx / a.band_rays : -1;
                    const unsigned peers = __match_any_sync(FULL, band);
                    if (was_finishing && lane == __ffs(peers) - 1) atomicAdd(a.band_done + band, __popc(peers));
                }
            }
            // ---------------- refill idle lanes ----------------
            if (!exhausted) {
                const unsigned idle = __ballot_sync(FULL, state == LANE_EMPTY);
                const int want = __popc(idle);
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.queue_head, (unsigned long long)want);
                base = __shfl_sync(FULL, base, 0);
                if (base + (unsigned long long)want >= (unsigned long long)q_len) exhausted = true;
                if (state == LANE_EMPTY) {
                    long long slot = (long long)base + __popc(idle & lt_mask);
                    bool take = slot < q_len;
                    if (use_hot && take) {
                        if (slot < n_hot) {
                            slot = (long long)__ldg(a.hot_list + slot);
                        } else {
                            slot -= n_hot;
                            take = !((__ldg(a.hot_mask + (slot >> 5)) >> (slot & 31)) & 1u);  // served from the list
                        }
                    }
                    if (take) {
                        idx = slot_to_ray(a, slot);
                        n_attempt = 0;
                        n_accept = 0;
                        rejected = false;
                        disk_hit = false;
                        pj = 0;
                        t = 0.0;
                        if constexpr (PREP) {
                            constexpr int PL = Prep<NK>::PL;
                            double v[2 * PL];
#pragma unroll
                            for (int pl = 0; pl < PL; pl++) {
                                const double2 w = __ldg(a.prep + (long long)pl * a.n + slot);
                                v[2 * pl] = w.x;
                                v[2 * pl + 1] = w.y;
                            }
                            const double h0 = Prep<NK>::unpack(v, k, x, K[0]);
                            h_abs = h0;
                            state = !(h0 < 0.0) ? LANE_RUNNING
                                                : (h0 == -1.0 ? (int)START_INSIDE_HOLE
                                                              : (h0 == -2.0 ? (int)STEP_FAILED : MISSED_SPHERE));
                        } else {
                            double x0[3], k0[3];
                            const bool enters = load_ray<IN>(a, idx, x0, k0);
                            if (a.coords && enters) iso_to_schw(a.rs, x0, k0);
                            if (!enters) {
                                state = MISSED_SPHERE;
                            } else if (!init_state<NK>(x0, k0, a.rs, a.r_hor, k, x)) {
                                state = START_INSIDE_HOLE;
                            } else if (!all_finite<NK>(k, x)) {
                                state = STEP_FAILED;  // singular entry (on the polar axis): scipy refuses such a y0
                            } else {
                                Rhs<NK>::eval(k, x, a.rs, K[0]);  // f0 (rk.py:94)
                                h_abs = initial_step<NK>(k, x, K[0], a.rs, a.rtol, a.atol, t_bound, a.max_step);
                                state = LANE_RUNNING;
                            }
                        }
                    }
                }
            }
            continue;
        }

        if (state == LANE_RUNNING) {
            // ---------------- one RK45 attempt (rk.py:111-176) ----------------
            // min_step = 10 ulp(t) <= 2.3e-15 t: the exact value is only formed when h_abs is that small
            bool too_small = false;
            if (a.has_max_step && !rejected && lt_nn(a.max_step, h_abs)) h_abs = a.max_step;
            if (lt_nn(h_abs, t * 2.3e-15)) {
                const double min_step = min_step_at(t);
                if (!rejected && lt_nn(h_abs, min_step)) h_abs = min_step;
                too_small = lt_nn(h_abs, min_step);
            }
            if (too_small) {
                state = STEP_FAILED;  // TOO_SMALL_STEP; (k, x) hold the last accepted state
            } else {
                const double t_new = min_nn(t + h_abs, t_bound);  // clip to t_bound (rk.py:139-140)
                const double h = t_new - t;  // >= 0: integration runs forward in lambda
                h_abs = h;
                n_attempt++;
                const double esum = rk45_attempt<NK>(k, x, K, kn, xn, h, a.rs, a.atol_over_rtol, a.inv_rtol2, lut);
                // 0.9 err^(-1/5), computed ONCE for the accepting and the rejecting lanes of the warp (a warp nearly
                // always holds both, and the inverse tenth root is ~35 cycles of the attempt)
                const double raw_factor = step_factor_raw<2 * NK>(esum);
                // esum = 2 NK (RMS error norm)^2: accepted iff error norm < 1 (rk.py:148)
                if (lt_nn(esum, 2.0 * NK)) {
                    n_accept++;
                    const double factor = min_nn(raw_factor, rejected ? 1.0 : 10.0);
                    // events on the accepted step (ivp.py:134-158).  Horizon (direction 0): a running ray always has
                    // r > r_hor (it starts there and stops at its first crossing), so "g0 >= 0 and g1 <= 0" is just
                    // r_new <= r_hor and the upward branch cannot occur.  Sphere (direction +1): g0 <= 0 and g1 >= 0.
                    const bool act_h = xn[IR] <= a.r_hor;
                    const bool act_e = a.has_outer && le_nn(x[IR], a.r_sphere) && le_nn(a.r_sphere, xn[IR]);
                    if (act_h || act_e) {
                        state = act_h ? (act_e ? PEND_HE : PEND_H) : PEND_E;
                        h_abs = h;  // keep the step length for the dense output
                    } else {
                        if constexpr (DISK) {  // non-terminal plane-crossing event on this accepted step
                            if (!disk_hit) {
                                const double half_pi = 1.57079632679489661923, inv_pi = 0.31830988618379067154;
                                if (floor((x[2] - half_pi) * inv_pi) != floor((xn[2] - half_pi) * inv_pi)) {
                                    // cheap radial reject before the (divergent) root search: within the step
                                    // r stays inside [min(r0,r1) - mg, max(r0,r1) + mg], mg = |h| max|k_r|
                                    const double mg = fabs(h) * fmax(fabs(k[1]), fabs(kn[1]));
                                    if (fmin(x[1], xn[1]) - mg <= a.disk_r_out && fmax(x[1], xn[1]) + mg >= a.disk_r_in)
                                        disk_hit = disk_crossing(a, idx, k, x, K, h, 1.0);
                                }
                            }
                        }
                        if constexpr (POLY) pj = emit_polyline(a, idx, pj, k, x, K, t, h, t_new);
                        h_abs *= factor;
                        t = t_new;
                        rejected = false;
#pragma unroll
                        for (int i = 0; i < NK; i++) {
                            k[i] = kn[i];
                            x[i] = xn[i];
                            K[0][i] = K[6][i];  // FSAL
                        }
                        if (le_nn(t_bound, t)) state = LAMBDA_EXHAUSTED;
                    }
                } else {
                    h_abs *= max_nn(raw_factor, 0.2);
                    rejected = true;
                }
            }
        }
    }
}

// ---- courier: delivers finished bands of a shard into the frame owner's memory while the trace kernel runs ----------
// A few resident blocks on SMs the trace kernel leaves free.  Local band j (band_rays consecutive rays of the compact
// shard, the last one possibly shorter) is complete when band_done[j] equals its ray count; it then moves as 16-byte
// vectors to frame band first_band + j * band_stride of the owner's buffers - contiguous hundreds of kilobytes, the
// shape NVLink likes, instead of 24-byte stores scattered over every SM's service phases.  Polling reads bypass L1.
struct CourierArgs {
    const double* src_pos; const double* src_dir; const int32_t* src_status;   // compact shard (local)
    double* dst_pos; double* dst_dir; int32_t* dst_status;                      // frame (owner's memory)
    const int* band_done;
    int* claimed;        // [bands] 0 / 1
    int* n_claimed;      // bands delivered so far
    long long m, band_rays, first_band, band_stride;
    int* error;          // set to 1 if a band never completed (the trace kernel failed): no hang
    long long wait_limit;  // microsecond-scale polls without progress before a block gives up (the sweep pass delivers
                           // what is left); <= 0: sweep pass - take whatever is complete and unclaimed, never wait
};

__device__ __forceinline__ void courier_copy(char* __restrict__ dst, const char* __restrict__ src, long long bytes,
                                             int tid, int nthreads) {
    // 16-byte vectors; all offsets / sizes are multiples of 16 (band_rays is a multiple of 4)
    const long long nvec = bytes >> 4;
    const int4* s = reinterpret_cast<const int4*>(src);
    int4* d = reinterpret_cast<int4*>(dst);
    long long i = tid;
    for (; i + 3LL * nthreads < nvec; i += 4LL * nthreads) {
        const int4 v0 = __ldcg(s + i), v1 = __ldcg(s + i + nthreads), v2 = __ldcg(s + i + 2LL * nthreads),
                   v3 = __ldcg(s + i + 3LL * nthreads);
        d[i] = v0; d[i + nthreads] = v1; d[i + 2LL * nthreads] = v2; d[i + 3LL * nthreads] = v3;
    }
    for (; i < nvec; i += nthreads) d[i] = __ldcg(s + i);
    const long long tail = bytes & 15;   // status of a partial last band
    if (tid < tail) dst[(nvec << 4) + tid] = src[(nvec << 4) + tid];
}

__global__ void __launch_bounds__(1024) courier_kernel(const CourierArgs a) {
    // Every block carries whole bands and takes ANY completed one (bands complete out of order: one photon-ring ray
    // holds its band back by a tenth of a millisecond): warp 0 scans the counters 32 bands at a time and claims a ready
    // band with a compare-and-swap, then all 1024 threads move it.
    const long long nb = (a.m + a.band_rays - 1) / a.band_rays;
    const int lane = threadIdx.x & 31;
    __shared__ long long s_band;
    long long idle = 0;
    while (true) {
        if (threadIdx.x < 32) {
            long long found = -1;
            for (long long base = 0; base < nb && found < 0; base += 32) {
                const long long j = base + lane;
                bool ready = false;
                if (j < nb && *(volatile const int*)(a.claimed + j) == 0) {
                    const long long cnt = (a.m - j * a.band_rays < a.band_rays) ? (a.m - j * a.band_rays) : a.band_rays;
                    ready = *(volatile const int*)(a.band_done + j) >= (int)cnt;
                }
                unsigned mask = __ballot_sync(0xffffffffu, ready);
                while (mask && found < 0) {
                    const int l = __ffs(mask) - 1;
                    int won = 0;
                    if (lane == l) won = atomicCAS(a.claimed + j, 0, 1) == 0;
                    won = __shfl_sync(0xffffffffu, won, l);
                    if (won) found = base + l;
                    mask &= mask - 1;
                }
            }
            if (found < 0 && *(volatile const int*)a.n_claimed >= (int)nb) found = -2;   // everything is taken: done
            if (lane == 0) s_band = found;
        }
        __syncthreads();
        const long long j = s_band;
        __syncthreads();
        if (j == -2) break;
        if (j < 0) {
            if (a.wait_limit <= 0) break;            // sweep pass: nothing complete is left unclaimed
            __nanosleep(1000);
            if (++idle > a.wait_limit) break;        // no progress (e.g. kernels are being run one at a time, as
            continue;                                // under a sanitizer): the sweep pass after the trace delivers
        }
        idle = 0;
        __threadfence();
        const long long lo = j * a.band_rays;
        const long long cnt = (a.m - lo < a.band_rays) ? (a.m - lo) : a.band_rays;
        const long long dst = (a.first_band + j * a.band_stride) * a.band_rays;
        courier_copy((char*)(a.dst_pos + 3 * dst), (const char*)(a.src_pos + 3 * lo), cnt * 24, threadIdx.x, blockDim.x);
        courier_copy((char*)(a.dst_dir + 3 * dst), (const char*)(a.src_dir + 3 * lo), cnt * 24, threadIdx.x, blockDim.x);
        courier_copy((char*)(a.dst_status + dst), (const char*)(a.src_status + lo), cnt * 4, threadIdx.x, blockDim.x);
        if (threadIdx.x == 0) atomicAdd(a.n_claimed, 1);
    }
    __threadfence_system();
}

// Standalone generator: entry positions / directions of n camera rays as AoS [n][3] (+ hit flag 0 / MISSED_SPHERE)
__global__ void generate_rays_kernel(const Camera cam, long long n, double* __restrict__ pos, double* __restrict__ dir,
                                     int32_t* __restrict__ hit) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double p[3], d[3];
        const bool ok = camera_ray(cam, i, p, d);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            pos[3 * i + c] = ok ? p[c] : __longlong_as_double(0x7ff8000000000000LL);
            dir[3 * i + c] = d[c];
        }
        if (hit) hit[i] = ok ? 0 : MISSED_SPHERE;
    }
}

}  // namespace bhg
