// Throughput of fp32 reductions to global memory (red.global.add): scalar vs .v2 vs .v4, coalesced rows as in the
// source-image scatter of the marching kernel.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_probe red_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void red1(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red2(float* p, float a, float b) { asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory"); }
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// every warp walks rows of a (rows x 1024) float image; per step it adds to `width` consecutive floats per lane
template <int MODE>
__global__ void probe(float* buf, int rows, int iters, int shift) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int it = 0; it < iters; ++it) {
        const int row = (warp + it * nwarps) % rows;
        float* r = buf + (size_t)row * 1024 + (((warp * 37 + it * 32) & 511) & ~3) + (MODE == 0 || MODE == 3 ? (warp & 3) : 0);
        if (MODE == 0) { red1(r + lane + shift, 1.f); red1(r + lane + 1 + shift, 1.f); red1(r + lane + 2 + shift, 1.f); red1(r + lane + 3 + shift, 1.f); }   // 4 scalar, 128 lane-ops
        if (MODE == 1) { float* q = r + ((2 * lane) & ~1); red2(q, 1.f, 1.f); red2(q + 64, 1.f, 1.f); }                   // 2 x v2: 128 floats
        if (MODE == 2) { float* q = r + ((4 * lane) & ~3); red4(q, 1.f, 1.f, 1.f, 1.f); }                                  // 1 x v4: 128 floats
        if (MODE == 3) { red1(r + lane + shift, 1.f); }                                                                    // 1 scalar: 32 floats
    }
}

#include <cstdlib>
int main(int argc, char** argv) {
    const int rows = argc > 1 ? atoi(argv[1]) : 8192;
    float* buf; cudaMalloc(&buf, sizeof(float) * rows * 1024); cudaMemset(buf, 0, sizeof(float) * rows * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 64, iters = 2000;
    const char* names[4] = {"4 x scalar (128 floats/warp-step)", "2 x v2     (128 floats/warp-step)", "1 x v4     (128 floats/warp-step)", "1 x scalar ( 32 floats/warp-step)"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<blocks, threads>>>(buf, rows, iters, 0);
            if (mode == 1) probe<1><<<blocks, threads>>>(buf, rows, iters, 0);
            if (mode == 2) probe<2><<<blocks, threads>>>(buf, rows, iters, 0);
            if (mode == 3) probe<3><<<blocks, threads>>>(buf, rows, iters, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double floats = (double)blocks * threads / 32 * iters * (mode == 3 ? 32 : 128);
        printf("%s: %.3f ms, %.1f G float-adds/s = %.1f per clock @1.965 GHz\n", names[mode], ms, floats / ms * 1e-6, floats / (ms * 1e-3) / 1.965e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
