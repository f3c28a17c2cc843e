// Does the register-file read bandwidth explain why non-FP64 instructions do not hide behind FP64 ones in the trace
// kernel?  8 DFMA per iteration with 1, 2 or 3 DISTINCT 64-bit register operands each (the others are loop-invariant
// and sit in the operand-reuse cache / are the accumulator itself), interleaved 1:1 with M/8 two-operand ALU
// instructions.  One 512-thread block per SM = 4 warps per scheduler.  Reported: cycles per warp-iteration per scheduler
// (16 = FP64-pipe floor).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o regread_bench scripts/regread_bench.cu && ./regread_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int NOPS, int KIND, int M>
__global__ void __launch_bounds__(512, 1) k(double* sink, long long* cycles, int iters, double m, unsigned key) {
    double a[8], b[8], d[8];
    unsigned u[8], v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
        b[i] = 1.0 + 1e-10 * (threadIdx.x + 3 * i);
        d[i] = 1e-12 * (threadIdx.x + 5 * i);
        u[i] = threadIdx.x * 2654435761u + i;
        v[i] = threadIdx.x * 40503u + 7 * i;
    }
    const double c = 1e-12;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (NOPS == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(c));
            if (NOPS == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(c));
            if (NOPS == 3) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(d[i]));
            if (NOPS == 4) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(a[i]) : "d"(b[(i + 1) & 7]), "d"(d[(i + 2) & 7]), "d"(b[(i + 5) & 7]));
#pragma unroll
            for (int j = 0; j < M; j++) {
                const int q = (i + 4 * j) & 7;
                if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[q]) : "r"(v[q]), "r"(key));
                if (KIND == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[q]) : "r"(v[q]), "r"(key));
                if (KIND == 2) asm volatile("mov.b32 %0, %1;" : "=r"(u[q]) : "r"(v[(q + 1) & 7]));
                if (KIND == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[q]) : "r"(v[q]));
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    double s = 0;
    unsigned x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += a[i] + b[i] + d[i]; x ^= u[i] ^ v[i]; }
    if (s == 12345.678 || x == 0x12345u) sink[threadIdx.x & 1023] = s + x;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NOPS, int KIND, int M>
double run(int sms, double* sink, long long* d_cyc, int iters) {
    k<NOPS, KIND, M><<<sms, 512>>>(sink, d_cyc, iters, 1.0000001, 3u);
    k<NOPS, KIND, M><<<sms, 512>>>(sink, d_cyc, iters, 1.0000001, 3u);
    cudaDeviceSynchronize();
    static long long h[1024];
    cudaMemcpy(h, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double tot = 0;
    for (int i = 0; i < sms; i++) tot += (double)h[i];
    return tot / sms / iters / 4.0;
}

template <int NOPS>
void sweep(int sms, double* sink, long long* d_cyc, int iters) {
    const char* kn[] = {"LOP3", "IMAD", "MOV", "IADD"};
    printf("DFMA with %d distinct register operand(s)%s:\n", NOPS == 4 ? 3 : NOPS, NOPS == 4 ? " (non-accumulating)" : "");
    printf("  %-5s +0 %6.2f | +8 %6.2f | +16 %6.2f\n", kn[0], run<NOPS, 0, 0>(sms, sink, d_cyc, iters), run<NOPS, 0, 1>(sms, sink, d_cyc, iters), run<NOPS, 0, 2>(sms, sink, d_cyc, iters));
    printf("  %-5s +0 %6.2f | +8 %6.2f | +16 %6.2f\n", kn[1], run<NOPS, 1, 0>(sms, sink, d_cyc, iters), run<NOPS, 1, 1>(sms, sink, d_cyc, iters), run<NOPS, 1, 2>(sms, sink, d_cyc, iters));
    printf("  %-5s +0 %6.2f | +8 %6.2f | +16 %6.2f\n", kn[2], run<NOPS, 2, 0>(sms, sink, d_cyc, iters), run<NOPS, 2, 1>(sms, sink, d_cyc, iters), run<NOPS, 2, 2>(sms, sink, d_cyc, iters));
    printf("  %-5s +0 %6.2f | +8 %6.2f | +16 %6.2f\n", kn[3], run<NOPS, 3, 0>(sms, sink, d_cyc, iters), run<NOPS, 3, 1>(sms, sink, d_cyc, iters), run<NOPS, 3, 2>(sms, sink, d_cyc, iters));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* sink;
    long long* d_cyc;
    cudaMalloc(&sink, 1024 * sizeof(double));
    cudaMalloc(&d_cyc, 1024 * sizeof(long long));
    const int iters = 20000;
    printf("# %s, %d SMs, 4 warps per scheduler, %d iterations of 8 DFMA interleaved with M ALU instructions; cycles per warp-iteration per scheduler\n", p.name, sms, iters);
    sweep<1>(sms, sink, d_cyc, iters);
    sweep<2>(sms, sink, d_cyc, iters);
    sweep<3>(sms, sink, d_cyc, iters);
    sweep<4>(sms, sink, d_cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    printf("# %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
