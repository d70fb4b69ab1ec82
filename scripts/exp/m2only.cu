// experiment TU: the single-warp marching kernel alone (fast compile for SASS inspection)
#include "../../monodepth2.jl_b200/csrc/md2_common.cuh"
#include "../../monodepth2.jl_b200/csrc/md2_fused.cuh"
#include "../../monodepth2.jl_b200/csrc/md2_march2.cuh"
#ifndef CC
#define CC 1
#endif
#ifndef SS
#define SS 2
#endif
#ifndef AMK
#define AMK false
#endif
namespace md2 {
__device__ __forceinline__ float warp_reduce_32(float (&v)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}
__global__ void __maxnreg__((March2<CC, SS, AMK>::MAXREG))
m2(const __grid_constant__ FusedParams p, int strips, int chunks, int q_full, int lgroups) {
    extern __shared__ __align__(16) float wsm[];
    using M = March2<CC, SS, AMK>;
    constexpr int NP = M::NPART;
    const int lane = threadIdx.x;
    const int ipg = strips * chunks;
    const int n_full = strips * q_full;
    const int ipi = n_full + (chunks > q_full ? lgroups : 0);
    const int items = ipi * p.L * p.N;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int z = it / ipi, rem = it - z * ipi;
        int cy, sx, sx_end;
        if (rem < n_full) { cy = rem / strips; sx = rem - cy * strips; sx_end = sx + 1; }
        else { cy = q_full; sx = (rem - n_full) * strips / lgroups; sx_end = (rem - n_full + 1) * strips / lgroups; }
        for (; sx < sx_end; ++sx) {
            float v[32];
            M::run(p, sx, cy, z, lane, wsm, v);
            const float tot = warp_reduce_32(v);
            if (lane == 0 || (lane >= NSTAT && lane < NP)) p.partial[((long long)z * ipg + cy * strips + sx) * NP + lane] = tot;
        }
    }
}
}
