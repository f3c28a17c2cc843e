// Microbenchmark behind the issue-slot model of the trace kernel: does an FP64-pipe instruction
// occupy the warp scheduler's issue port for both cycles of its 2-cycle pipe occupancy (16 FP64 lanes per SM
// sub-partition), or can another pipe's instruction issue in the shadow cycle?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_model_bench scripts/issue_model_bench.cu && ./issue_model_bench
//
// Every thread runs 8 independent DFMA chains; per group of 8 DFMAs it also issues M independent instructions of
// another kind (IMAD, LOP3, SEL, FMUL, MUFU, LDC, PRMT).  One 512-thread block per SM = 4 warps per scheduler, the
// trace kernel's residency.  Reported: SM cycles per warp-iteration per scheduler,
//   = 16 if only the FP64 pipe matters (8 DFMA x 2 cycles),
//   = 16 + M if an FP64 instruction blocks the issue port for two cycles and everything else costs one more.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

enum Kind { K_NONE = 0, K_IMAD, K_LOP3, K_SEL, K_FMUL, K_MUFU, K_LDC, K_PRMT, K_DADD, K_DMUL, K_IADD, K_LAST };
static const char* kind_name[] = {"none", "IMAD", "LOP3", "SEL", "FMUL", "MUFU.EX2", "LDC", "PRMT", "DADD", "DMUL", "IADD3"};

__constant__ unsigned c_words[64];

template <int KIND, int M>
__global__ void __launch_bounds__(512, 1) mix_kernel(double* sink, long long* cycles, int iters, double m, unsigned key) {
    double a[8];
    unsigned b[8];
    float f[8];
    double e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
        b[i] = threadIdx.x * 2654435761u + i;
        f[i] = 1.0f + 1e-3f * i;
        e[i] = 1.0 + 1e-7 * i;
    }
    const double c = 1e-12;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
#pragma unroll
        for (int j = 0; j < M; j++) {
            const int i = j & 7;
            if (KIND == K_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(key), "r"(key));
            if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(key), "r"(b[(i + 1) & 7]));
            if (KIND == K_SEL) asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0; selp.b32 %0, %0, %2, p; }" : "+r"(b[i]) : "r"(key), "r"(b[(i + 3) & 7]));
            if (KIND == K_FMUL) asm volatile("mul.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(1.0000001f));
            if (KIND == K_MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            if (KIND == K_LDC) asm volatile("ld.const.u32 %0, [%1];" : "=r"(b[i]) : "l"((const void*)(c_words + (b[i] & 63))));
            if (KIND == K_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x1032;" : "+r"(b[i]) : "r"(key));
            if (KIND == K_DADD) asm volatile("add.f64 %0, %0, %1;" : "+d"(e[i]) : "d"(c));
            if (KIND == K_DMUL) asm volatile("mul.f64 %0, %0, %1;" : "+d"(e[i]) : "d"(m));
            if (KIND == K_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(key));
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    double s = 0;
    unsigned u = 0;
    float g = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += a[i] + e[i]; u ^= b[i]; g += f[i]; }
    if (s == 12345.678 || u == 0x12345u || g == 3.25f) sink[threadIdx.x & 1023] = s + u + g;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND, int M>
double run(int sms, double* sink, long long* d_cyc, int iters) {
    mix_kernel<KIND, M><<<sms, 512>>>(sink, d_cyc, iters, 1.0000001, 3u);   // warm-up
    mix_kernel<KIND, M><<<sms, 512>>>(sink, d_cyc, iters, 1.0000001, 3u);
    cudaDeviceSynchronize();
    static long long h[1024];
    cudaMemcpy(h, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double tot = 0;
    for (int i = 0; i < sms; i++) tot += (double)h[i];
    return tot / sms / iters / 4.0;  // cycles per warp-iteration per scheduler (4 warps share one scheduler)
}

template <int KIND>
void sweep(int sms, double* sink, long long* d_cyc, int iters) {
    const double r0 = run<KIND, 0>(sms, sink, d_cyc, iters), r2 = run<KIND, 2>(sms, sink, d_cyc, iters),
                 r4 = run<KIND, 4>(sms, sink, d_cyc, iters), r8 = run<KIND, 8>(sms, sink, d_cyc, iters),
                 r16 = run<KIND, 16>(sms, sink, d_cyc, iters);
    printf("%-9s per 8 DFMA: +0 %6.2f | +2 %6.2f | +4 %6.2f | +8 %6.2f | +16 %6.2f   cycles/warp-iteration  "
           "(marginal cost per extra instruction: %.2f / %.2f / %.2f / %.2f cycles)\n",
           kind_name[KIND], r0, r2, r4, r8, r16, (r2 - r0) / 2, (r4 - r0) / 4, (r8 - r0) / 8, (r16 - r0) / 16);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* sink;
    long long* d_cyc;
    cudaMalloc(&sink, 1024 * sizeof(double));
    cudaMalloc(&d_cyc, 1024 * sizeof(long long));
    unsigned hw[64];
    for (int i = 0; i < 64; i++) hw[i] = i * 7u + 1u;
    cudaMemcpyToSymbol(c_words, hw, sizeof(hw));
    const int iters = 20000;
    printf("# %s, %d SMs, one 512-thread block per SM (4 warps per scheduler), %d iterations of 8 independent DFMA + M others\n",
           p.name, sms, iters);
    sweep<K_IMAD>(sms, sink, d_cyc, iters);
    sweep<K_IADD>(sms, sink, d_cyc, iters);
    sweep<K_LOP3>(sms, sink, d_cyc, iters);
    sweep<K_SEL>(sms, sink, d_cyc, iters);
    sweep<K_PRMT>(sms, sink, d_cyc, iters);
    sweep<K_FMUL>(sms, sink, d_cyc, iters);
    sweep<K_MUFU>(sms, sink, d_cyc, iters);
    sweep<K_LDC>(sms, sink, d_cyc, iters);
    sweep<K_DADD>(sms, sink, d_cyc, iters);
    sweep<K_DMUL>(sms, sink, d_cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    printf("# %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
