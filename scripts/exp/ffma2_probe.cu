// experiment: issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, alone and
// mixed with integer ALU work.  Prints warp-instructions per cycle per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float s) {
    float a[8]; u64 p[8]; int q[8];
    for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * 0.001f + k; p[k] = pk(a[k], a[k] + 1.f); q[k] = threadIdx.x + k; }
    const u64 s2 = pk(s, s), t2 = pk(0.5f, 0.25f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) { a[k] = fmaf(a[k], s, 0.5f); }                      // 8 FFMA
            if (MODE == 1) { FMA2(p[k], p[k], s2, t2); }                         // 8 FFMA2
            if (MODE == 2) { a[k] = fmaf(a[k], s, 0.5f); q[k] = (q[k] ^ i) + k; }   // 8 FFMA + 8x2 ALU
            if (MODE == 3) { FMA2(p[k], p[k], s2, t2); q[k] = (q[k] ^ i) + k; }
            if (MODE == 4) { a[k] = fmaf(a[k], s, a[(k + 1) & 7]); }             // 3-register FFMA
            if (MODE == 5) { FMA2(p[k], p[k], s2, p[(k + 1) & 7]); }
        }
    }
    float r = 0.f;
    for (int k = 0; k < 8; ++k) r += a[k] + __uint_as_float((unsigned)p[k]) + __uint_as_float((unsigned)(p[k] >> 32)) + q[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, int fp_per_it, int alu_per_it) {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<148 * 8, 256>>>(d, 100, 1.0001f);
    cudaEventRecord(e0);
    probe<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double winst = (double)iters * (fp_per_it + alu_per_it) * 8 * 8 /*warps per SM / 4 SMSP... */;
    // per SMSP: 8 blocks x 8 warps / 4 = 16 warps
    const double per_smsp = (double)iters * (fp_per_it + alu_per_it) * 16;
    printf("%-34s %8.3f ms  %.3f warp-inst/clk/SMSP (fp %d + alu %d per iter; nominal clock %d kHz)\n", name, ms, per_smsp / cycles, fp_per_it, alu_per_it, clk);
    (void)winst; cudaFree(d);
}
int main() {
    run<0>("FFMA imm", 8, 0);
    run<1>("FFMA2", 8, 0);
    run<2>("FFMA + 2 ALU", 8, 16);
    run<3>("FFMA2 + 2 ALU", 8, 16);
    run<4>("FFMA 3-reg", 8, 0);
    run<5>("FFMA2 3-reg", 8, 0);
    return 0;
}
