// Device-side primary-ray generation and sphere entry: the caller side of the hot path.
//
// Restates the reference's pinhole generator
//   /root/reference/raytracer/RelativisticRenderEngine.py:185-189 (aspect, dx, dy), :195-230 (s -> y -> x order,
//   direction = R_cam . (x_render + jitter_x, y_render + jitter_y, -1), normalised)
// and the sphere-of-influence entry of /root/reference/raytracer/LimitedRelativisticRenderEngine.py:224,265
// (flat ray_cast hit on the "isBH" sphere minus its centre), with a counter-based Philox-4x32-10 jitter stream
// (bit-identical to raygen.philox4x32_10 on the host) instead of the reference's sequential Mersenne Twister,
// so any ray can be generated independently from its index.
#pragma once
#include <cstdint>

namespace bhg {

struct Camera {
    double origin[3];    // camera position relative to the black-hole centre
    double rot[9];       // row-major 3x3 camera-to-world rotation (camera looks along local -z, +y is up)
    double fov_x, fov_y;
    double r_sphere;
    long long first_ray; // offset of ray 0 of this call in the frame's s -> y -> x order
    unsigned long long seed;
    int width, height;
    int jitter;          // 0: pixel centres (u = 0.5), 1: Philox
};

__device__ __forceinline__ void philox4x32_10(unsigned long long ctr, unsigned long long seed, unsigned (&out)[4]) {
    unsigned c0 = (unsigned)ctr, c1 = (unsigned)(ctr >> 32), c2 = 0u, c3 = 0u;
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) from two 32-bit words (the construction of MT19937's random())
__device__ __forceinline__ double u01(unsigned hi, unsigned lo) {
    const double a = (double)(hi >> 5), b = (double)(lo >> 6);
    return __dmul_rn(__dadd_rn(__dmul_rn(a, 67108864.0), b), 1.0 / 9007199254740992.0);
}

// Ray `idx` of the call: unit direction d and first intersection p with the sphere (relative to its centre).
// Returns false when the flat ray misses the sphere (p is then undefined).
// Non-contracted arithmetic (__dmul_rn / __dadd_rn) mirrors the host generator's operation order.
__device__ __forceinline__ bool camera_ray(const Camera& c, long long idx, double (&p)[3], double (&d)[3]) {
    const long long g = c.first_ray + idx;
    const int px = (int)(g % c.width);
    const int py = (int)((g / c.width) % c.height);
    double u1 = 0.5, u2 = 0.5;
    if (c.jitter) {
        unsigned r[4];
        philox4x32_10((unsigned long long)g, c.seed, r);
        u1 = u01(r[0], r[1]);
        u2 = u01(r[2], r[3]);
    }
    const double aspect = (double)c.height / (double)c.width;
    const double dx = 1.0 / (double)c.width, dy = aspect / (double)c.height;
    const double xr = __dadd_rn(__dmul_rn(c.fov_x, (double)(px - c.width / 2)) / (double)c.width,
                                __dmul_rn(dx, __dadd_rn(u1, -0.5)));
    const double yr = __dadd_rn(__dmul_rn(__dmul_rn(c.fov_y, (double)(py - c.height / 2)) / (double)c.height, aspect),
                                __dmul_rn(dy, __dadd_rn(u2, -0.5)));
    double v[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
        v[i] = __dadd_rn(__dadd_rn(__dmul_rn(c.rot[3 * i], xr), __dmul_rn(c.rot[3 * i + 1], yr)), -c.rot[3 * i + 2]);
    const double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2])));
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = v[i] / nrm;
    // first intersection with |p| = r_sphere
    const double od = __dadd_rn(__dadd_rn(__dmul_rn(d[0], c.origin[0]), __dmul_rn(d[1], c.origin[1])),
                                __dmul_rn(d[2], c.origin[2]));
    const double oo = __dadd_rn(__dadd_rn(__dmul_rn(c.origin[0], c.origin[0]), __dmul_rn(c.origin[1], c.origin[1])),
                                __dmul_rn(c.origin[2], c.origin[2]));
    // camera inside the sphere of influence (the RRE / CAM engines put it there: RelativisticRenderEngine.py:278,
    // RelativisticRenderEngineCamEdition.py:212): the ray starts at the camera itself
    if (oo < __dmul_rn(c.r_sphere, c.r_sphere)) {
#pragma unroll
        for (int i = 0; i < 3; i++) p[i] = c.origin[i];
        return true;
    }
    const double disc = __dadd_rn(__dmul_rn(od, od), -__dadd_rn(oo, -__dmul_rn(c.r_sphere, c.r_sphere)));
    if (!(disc >= 0.0) || !(od < 0.0)) return false;
    const double s = __dadd_rn(-od, -sqrt(disc));
    if (!(s >= 0.0)) return false;
#pragma unroll
    for (int i = 0; i < 3; i++) p[i] = __dadd_rn(c.origin[i], __dmul_rn(s, d[i]));
    return true;
}

}  // namespace bhg
