// Second register-read probe: which instructions hide in the second cycle of a DFMA, as a function of the DFMA's
// distinct register operands.  Same set-up as regread_bench.cu (8 DFMA per iteration, one 512-thread block per SM).
#include <cstdio>
#include <cuda_runtime.h>

template <int NOPS, int KIND, int M>
__global__ void __launch_bounds__(512, 1) k(double* sink, long long* cycles, int iters, double m, unsigned key) {
    double a[8], b[8], d[8];
    unsigned u[8], v[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
        b[i] = 1.0 + 1e-10 * (threadIdx.x + 3 * i);
        d[i] = 1e-12 * (threadIdx.x + 5 * i);
        u[i] = threadIdx.x * 2654435761u + i;
        v[i] = threadIdx.x * 40503u + 7 * i;
        f[i] = 1.0f + i;
    }
    const double c = 1e-12;
    double mreg = m * 1.0000001;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (NOPS == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(c));
            if (NOPS == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(c));
            if (NOPS == 3) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(d[i]));
            if (NOPS == 5) asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(a[i]) : "d"(mreg), "d"(d[i]));   // invariant reg (reuse) + 2
            if (NOPS == 6) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[i]));                  // DMUL 2 regs
            if (NOPS == 7) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(m));                     // DMUL 1 reg
#pragma unroll
            for (int j = 0; j < M; j++) {
                const int q = (i + 4 * j) & 7;
                if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[q]) : "r"(v[q]), "r"(key));   // 2-3 reads
                if (KIND == 1) asm volatile("and.b32 %0, %0, 0x7fffffff;" : "+r"(u[q]));                            // 1 read
                if (KIND == 2) asm volatile("mov.b32 %0, %1;" : "=r"(u[q]) : "r"(v[(q + 1) & 7]));                  // 1 read
                if (KIND == 3) asm volatile("mov.b32 %0, 0x1234;" : "=r"(u[q]));                                    // 0 reads
                if (KIND == 4) asm volatile("{ .reg .pred p; setp.lt.u32 p, %1, %2; selp.b32 %0, %0, %1, p; }" : "+r"(u[q]) : "r"(v[q]), "r"(key));  // ISETP + SEL
                if (KIND == 5) asm volatile("add.u32 %0, %0, 0x11;" : "+r"(u[q]));                                  // IADD imm: 1 read
                if (KIND == 6) asm volatile("mul.f32 %0, %0, 0f3F800001;" : "+f"(f[q]));                             // FMUL 1 read
                if (KIND == 7) asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[q]));                                     // shift 1 read
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    double s = 0;
    unsigned x = 0;
    float g = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += a[i] + b[i] + d[i]; x ^= u[i] ^ v[i]; g += f[i]; }
    if (s == 12345.678 || x == 0x12345u || g == 1.5f) sink[threadIdx.x & 1023] = s + x + g;
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
void sweep(const char* name, int sms, double* sink, long long* d_cyc, int iters) {
    printf("%-34s alone %6.2f | +8: LOP3(2r) %6.2f  AND(1r) %6.2f  MOV(1r) %6.2f  MOV(imm) %6.2f  ISETP+SEL %6.2f  IADD(1r) %6.2f  FMUL(1r) %6.2f  SHL(1r) %6.2f\n",
           name, run<NOPS, 0, 0>(sms, sink, d_cyc, iters), run<NOPS, 0, 1>(sms, sink, d_cyc, iters),
           run<NOPS, 1, 1>(sms, sink, d_cyc, iters), run<NOPS, 2, 1>(sms, sink, d_cyc, iters),
           run<NOPS, 3, 1>(sms, sink, d_cyc, iters), run<NOPS, 4, 1>(sms, sink, d_cyc, iters),
           run<NOPS, 5, 1>(sms, sink, d_cyc, iters), run<NOPS, 6, 1>(sms, sink, d_cyc, iters),
           run<NOPS, 7, 1>(sms, sink, d_cyc, iters));
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
    printf("# %s: cycles per warp-iteration per scheduler, 8 FP64 instructions per iteration interleaved 1:1 with 8 others\n", p.name);
    sweep<1>("DFMA acc,UR,UR (1 reg)", sms, sink, d_cyc, iters);
    sweep<2>("DFMA acc,reg,UR (2 regs)", sms, sink, d_cyc, iters);
    sweep<3>("DFMA acc,reg,reg (3 regs)", sms, sink, d_cyc, iters);
    sweep<5>("DFMA inv,acc,reg (3 regs, 1 invariant)", sms, sink, d_cyc, iters);
    sweep<6>("DMUL acc,reg (2 regs)", sms, sink, d_cyc, iters);
    sweep<7>("DMUL acc,UR (1 reg)", sms, sink, d_cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    printf("# %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
