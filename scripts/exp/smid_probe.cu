// experiment: how does the block scheduler place a grid smaller than the resident capacity?
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(64) probe(int* smid, long long* t0, int spin) {
    extern __shared__ float sm[];
    unsigned id; asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    long long t = clock64();
    if (threadIdx.x == 0) { smid[blockIdx.x] = id; t0[blockIdx.x] = t; }
    float x = threadIdx.x;
    for (int i = 0; i < spin; ++i) x = x * 1.0001f + 0.5f;
    if (x == 123.f) sm[0] = x;
}
int main() {
    for (int blocks : {960, 1184, 1440, 1920}) {
        int* d; long long* t; cudaMalloc(&d, blocks * 4); cudaMalloc(&t, blocks * 8);
        size_t smem = 26 * 1024;   // 8 blocks/SM by shared memory
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe<<<blocks, 64, smem>>>(d, t, 200000);
        cudaDeviceSynchronize();
        std::vector<int> h(blocks); cudaMemcpy(h.data(), d, blocks * 4, cudaMemcpyDeviceToHost);
        int cnt[256] = {0}; int mx = 0;
        for (int b = 0; b < blocks; ++b) { cnt[h[b]]++; if (h[b] > mx) mx = h[b]; }
        int hist[64] = {0}; int used = 0;
        for (int s = 0; s <= mx; ++s) { hist[cnt[s]]++; if (cnt[s]) used++; }
        printf("blocks %d: max smid %d, SMs used %d; blocks/SM histogram:", blocks, mx, used);
        for (int k = 0; k < 64; ++k) if (hist[k]) printf(" %d:%d", k, hist[k]);
        printf("\n  first 24 blocks -> smid:");
        for (int b = 0; b < 24; ++b) printf(" %d", h[b]);
        printf("\n");
        cudaFree(d); cudaFree(t);
    }
    return 0;
}
