// Auxiliary kernels: counter reduction (roofline accounting), FP64 self-test, DFMA peak microbenchmark.
#pragma once
#include "geodesic_core.cuh"

namespace bhg {

// totals[0] = sum attempts, totals[1] = sum accepted, totals[2] = rays that were integrated (status != START_INSIDE)
__global__ void sum_counters_kernel(const int32_t* __restrict__ counters, const int32_t* __restrict__ status,
                                    long long n, long long* __restrict__ totals) {
    long long a = 0, b = 0, c = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        a += counters[i];
        b += counters[n + i];
        c += status[i] != START_INSIDE_HOLE;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long*)&totals[0], (unsigned long long)a);
        atomicAdd((unsigned long long*)&totals[1], (unsigned long long)b);
        atomicAdd((unsigned long long*)&totals[2], (unsigned long long)c);
    }
}

__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* p = (unsigned long long*)addr;
    unsigned long long old = *p, assumed;
    do {
        assumed = old;
        if (__longlong_as_double(assumed) >= v) break;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}

// out[0]: max rel. error of fast_rcp vs IEEE 1/a ; out[1]: inv_tenth_root vs pow(a,-0.1) ;
// out[2]: RHS (reciprocal form) vs the textbook division form ; out[3]: 3-DFMA reciprocal ;
// out[4]: table sincos vs library sincos (absolute) ; out[5]: fifth_root vs pow(a, 0.2)
__global__ void selftest_kernel(double* out) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    double e_rcp = 0, e_root = 0, e_rhs = 0, e_rcp3 = 0, e_sc = 0, e_root5 = 0;
    for (int j = gid; j < nth * 16; j += nth) {
        // log-uniform-ish positive and negative operands
        const double u = (j + 0.5) / (nth * 16.0);
        const double a = exp((u - 0.5) * 80.0) * ((j & 1) ? -1.0 : 1.0);
        const double r0 = 1.0 / a, r1 = fast_rcp5(a);
        e_rcp = fmax(e_rcp, fabs(r1 - r0) / fabs(r0));
        e_rcp3 = fmax(e_rcp3, fabs(fast_rcp(a) - r0) / fabs(r0));
        {
            // table sincos vs the library over |x| < 40
            const double xs = (u - 0.5) * 80.0;
            double s0, c0, s1, c1;
            sincos(xs, &s0, &c0);
            sincos_tab(xs, &s1, &c1);
            e_sc = fmax(e_sc, fmax(fabs(s1 - s0), fabs(c1 - c0)));
        }
        const double b = exp((u - 0.5) * 40.0);  // 2e-9 .. 5e8 covers [1e-12,1e8] core range partially
        if (b > 1e-12 && b < 1e8) {
            const double p0 = pow(b, -0.1), p1 = inv_tenth_root(b);
            e_root = fmax(e_root, fabs(p1 - p0) / p0);
            const double w0 = pow(b, 0.2), w1 = fifth_root(b);
            e_root5 = fmax(e_root5, fabs(w1 - w0) / w0);
        }
        // RHS comparison at a generic state
        const double r = 2.2 + 60.0 * u, th = 0.05 + 3.0 * u, rs = 2.0;
        double y[8] = {1.0 + u, 0.0, 0.9 - 1.8 * u, r, 0.01 * (u - 0.3), th, 0.02 * (0.7 - u), 1.0};
        double kk[4] = {y[0], y[2], y[4], y[6]}, xx[4] = {y[1], y[3], y[5], y[7]}, ff[4];
        Rhs<4>::eval(kk, xx, rs, ff);
        double f[8] = {ff[0], 0, ff[1], 0, ff[2], 0, ff[3], 0};
        const double s = sin(th), rm = r - rs;
        const double g0 = -y[2] * y[0] * rs / (r * rm);
        const double g2 = (y[2] * y[2] * r * r * rs - y[0] * y[0] * rs * rm * rm +
                           2 * r * r * r * rm * rm * (y[6] * y[6] * s * s + y[4] * y[4])) / (2 * r * r * r * rm);
        const double g4 = y[6] * y[6] * sin(2 * th) / 2 - 2 * y[2] * y[4] / r;
        const double g6 = -2 * y[6] * (y[2] + y[4] * r / tan(th)) / r;
        const double sc = fabs(g0) + fabs(g2) + fabs(g4) + fabs(g6) + 1e-300;
        e_rhs = fmax(e_rhs, (fabs(f[0] - g0) + fabs(f[2] - g2) + fabs(f[4] - g4) + fabs(f[6] - g6)) / sc);
    }
    atomic_max_double(&out[0], e_rcp);
    atomic_max_double(&out[1], e_root);
    atomic_max_double(&out[2], e_rhs);
    atomic_max_double(&out[3], e_rcp3);
    atomic_max_double(&out[4], e_sc);
    atomic_max_double(&out[5], e_root5);
}

// Sky-lookup coordinates of the exit directions: the equirectangular mapping of the
// reference's background_hit (raytracer/RelativisticRenderEngine.py:366-378,
// raytracer/LimitedRelativisticRenderEngine.py:383-408):
//     theta = 1 - acos(d_z)/pi ; phi = atan2(d_y, d_x)/pi ; texture.evaluate((-phi, 2 theta - 1, 0))
// -> uv = (-phi, 2 theta - 1), computed in FP64 and stored as float2 (texture coordinates).  Captured /
// failed rays get NaN (the reference paints them black without a lookup, RRE.py:242-244).
__global__ void sky_uv_kernel(const double* __restrict__ dir, const int32_t* __restrict__ status, long long n,
                              float2* __restrict__ uv) {
    const double inv_pi = 0.31830988618379067154;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int st = status ? status[i] : ESCAPED;
        float2 o;
        if (st == CAPTURED || st == START_INSIDE_HOLE || st == STEP_FAILED) {
            o.x = o.y = __int_as_float(0x7fc00000);
        } else {
            const double dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
            const double theta = 1.0 - acos(dz) * inv_pi;
            const double phi = atan2(dy, dx) * inv_pi;
            o.x = (float)(-phi);
            o.y = (float)(2.0 * theta - 1.0);
        }
        uv[i] = o;
    }
}

// 16 independent DFMA chains per thread; flops = 2 * 16 * iters per thread
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* sink, int iters, double m) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double c = 1e-12;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 12345.678) sink[threadIdx.x & 1023] = s;
}

}  // namespace bhg
