// Third probe: is the 3-cycle cost of a DFMA with three distinct register operands a register-bank artefact?
// Several operand patterns (different register triples after allocation); the SASS register numbers are printed by
// `cuobjdump -sass` next to the timing.  4 warps per scheduler, 8 DFMA per iteration.
#include <cstdio>
#include <cuda_runtime.h>

template <int PAT>
__global__ void __launch_bounds__(512, 1) k(double* sink, long long* cycles, int iters) {
    double a[8], b[8], d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
        b[i] = 1.0 + 1e-10 * (threadIdx.x + 3 * i);
        d[i] = 1e-12 * (threadIdx.x + 5 * i);
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (PAT == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(d[i]));
            if (PAT == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[(i + 1) & 7]), "d"(d[(i + 3) & 7]));
            if (PAT == 2) asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(a[i]) : "d"(b[(i + 2) & 7]), "d"(d[(i + 5) & 7]));
            if (PAT == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b[i]), "d"(d[(i + 1) & 7]));
            if (PAT == 4) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b[i]), "d"(b[(i + 1) & 7]));
            if (PAT == 5) asm volatile("fma.rn.f64 %0, %1, %1, %0;" : "+d"(a[i]) : "d"(b[i]));                      // 2 distinct
            if (PAT == 6) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(a[i]) : "d"(b[i]), "d"(d[i]), "d"(a[(i + 4) & 7]));  // dest != src
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + b[i] + d[i];
    if (s == 12345.678) sink[threadIdx.x & 1023] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int PAT>
void run(int sms, double* sink, long long* d_cyc, int iters) {
    k<PAT><<<sms, 512>>>(sink, d_cyc, iters);
    k<PAT><<<sms, 512>>>(sink, d_cyc, iters);
    cudaDeviceSynchronize();
    static long long h[1024];
    cudaMemcpy(h, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double tot = 0;
    for (int i = 0; i < sms; i++) tot += (double)h[i];
    printf("pattern %d: %.2f cycles per DFMA per scheduler-warp\n", PAT, tot / sms / iters / 4.0 / 8.0);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* sink;
    long long* d_cyc;
    cudaMalloc(&sink, 1024 * sizeof(double));
    cudaMalloc(&d_cyc, 1024 * sizeof(long long));
    run<0>(sms, sink, d_cyc, 20000); run<1>(sms, sink, d_cyc, 20000); run<2>(sms, sink, d_cyc, 20000);
    run<3>(sms, sink, d_cyc, 20000); run<4>(sms, sink, d_cyc, 20000); run<5>(sms, sink, d_cyc, 20000);
    run<6>(sms, sink, d_cyc, 20000);
    return cudaDeviceSynchronize() != cudaSuccess;
}
