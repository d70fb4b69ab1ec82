// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or shipped with the product.
// Runs the marching-warp kernel source (monodepth2.jl_b200/csrc/md2_march.cuh) on the CPU, every warp as
// 32 lockstep fibers (warp_emu.h), plus the host versions of the prep / adjoint steps around it, so that
// the halo / reflect / rolling-window / scatter logic of the fused CUDA kernel can be checked against the
// oracle in a container without a GPU.
// All pointers in the descriptor are HOST pointers here.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/md2.h"
#include "../../monodepth2.jl_b200/csrc/md2_fused.cuh"
#define MD2_WARP_EMU 1
#include "warp_emu.h"
#include "../../monodepth2.jl_b200/csrc/md2_march.cuh"
#include "../../monodepth2.jl_b200/csrc/md2_march2.cuh"

using namespace md2;

// marching-warp kernel (md2_march.cuh): a block of one (forward-only) or two (forward + backward
// producer/consumer pair) warps of lockstep fibers walks all work items of one (scale, image), like
// a persistent block of the CUDA kernel does
template <int C, int S, bool BWD>
static void run_march(const FusedParams& p, std::vector<double>& sums) {
    using M = March<C, S, BWD>;
    const int NP = M::NPART;
    const int strips = (p.W + M::OW - 1) / M::OW, chunks = (p.H + p.m_R - 1) / p.m_R;
    sums.assign((size_t)p.L * p.N * NP, 0.0);
    std::vector<float> wsm(M::SMEM_FLOATS + 4);
    float* wsm_al = (float*)(((uintptr_t)wsm.data() + 15) & ~(uintptr_t)15);
    WarpEmu emu;
    std::vector<float> lane_v((size_t)M::THREADS * 32);
    for (int z = 0; z < p.L * p.N; ++z) {
        for (int k = 0; k < M::SMEM_FLOATS; ++k) wsm_al[k] = NAN;   // poison: catches reads of never-written slots
        double* su = sums.data() + (size_t)z * NP;
        emu.run(M::THREADS, [&](int tid) {
            const int warp = tid >> 5, lane = tid & 31;
            int gslot = 0;
            if (tid == 0)
                for (int b = 0; b < MARCH_NBAR; ++b) emu_mb_init(reinterpret_cast<mbar_t*>(wsm_al + M::RING_FLOATS), b);
            if (BWD) emu_bar(0, 64, 1); else emu_ballot(0);
            for (int cy = 0; cy < chunks; ++cy)
                for (int sx = 0; sx < strips; ++sx) {
                    float v[32];
                    if (warp == 0) M::run_forward(p, sx, cy, z, lane, wsm_al, gslot, v);
                    else M::run_backward(p, sx, cy, z, lane, wsm_al, gslot, v);
                    for (int k = 0; k < 32; ++k) lane_v[(size_t)tid * 32 + k] = v[k];
                    if (BWD) emu_bar(0, 64, 1); else emu_ballot(0);
                    if (tid == 0)
                        for (int k = 0; k < NP; ++k) {   // float over the lanes of a warp (device: warp reduction), double across items (finish kernel)
                            for (int w = 0; w < M::THREADS / 32; ++w) {
                                float part = 0.f;
                                for (int t = 0; t < 32; ++t) part += lane_v[(size_t)(w * 32 + t) * 32 + k];
                                su[k] += (double)part;
                            }
                        }
                    if (BWD) emu_bar(0, 64, 1); else emu_ballot(0);
                }
        });
    }
}

// single-warp marching kernel (md2_march2.cuh): one warp of lockstep fibers walks all work items of a (scale, image)
template <int C, int S, bool AM, bool DBG = false, bool GRAD = true>
static void run_march2(const FusedParams& p, std::vector<double>& sums) {
    using M = March2<C, S, AM, DBG, GRAD>;
    const int NP = M::NPART;
    const int strips = (p.W + M::OW - 1) / M::OW, chunks = (p.H + p.m_R - 1) / p.m_R;
    sums.assign((size_t)p.L * p.N * NP, 0.0);
    std::vector<float> wsm(M::SMEM_FLOATS + 4);
    float* wsm_al = (float*)(((uintptr_t)wsm.data() + 15) & ~(uintptr_t)15);
    WarpEmu emu;
    std::vector<float> lane_v((size_t)32 * 32);
    for (int z = 0; z < p.L * p.N; ++z) {
        for (int k = 0; k < M::SMEM_FLOATS; ++k) wsm_al[k] = NAN;   // poison: catches reads of never-written slots
        double* su = sums.data() + (size_t)z * NP;
        emu.run(32, [&](int tid) {
            const int lane = tid & 31;
            for (int cy = 0; cy < chunks; ++cy)
                for (int sx = 0; sx < strips; ++sx) {
                    float v[32];
                    M::run(p, sx, cy, z, lane, wsm_al, v);
                    for (int k = 0; k < 32; ++k) lane_v[(size_t)tid * 32 + k] = v[k];
                    emu_ballot(0);
                    if (tid == 0)
                        for (int k = 0; k < NP; ++k) {   // float over the lanes (device: warp reduction), double across items (finish kernel)
                            float part = 0.f;
                            for (int t = 0; t < 32; ++t) part += lane_v[(size_t)t * 32 + k];
                            su[k] += (double)part;
                        }
                    emu_ballot(0);
                }
        });
    }
}

template <bool BWD>
static int dispatch(int C, int S, const FusedParams& p, std::vector<double>& sums, int variant) {
    if (variant == 1) {
        if (C == 1 && S == 1) { run_march<1, 1, BWD>(p, sums); return 0; }
        if (C == 1 && S == 2) { run_march<1, 2, BWD>(p, sums); return 0; }
        if (C == 3 && S == 1) { run_march<3, 1, BWD>(p, sums); return 0; }
        if (C == 3 && S == 2) { run_march<3, 2, BWD>(p, sums); return 0; }
        return 1;
    }
    if (variant == 2 && !BWD) {   // forward-only instantiations of the single-warp kernel
        const bool am = p.automask != nullptr;
        if (C == 1 && S == 1) { if (am) run_march2<1, 1, true, false, false>(p, sums); else run_march2<1, 1, false, false, false>(p, sums); return 0; }
        if (C == 1 && S == 2) { if (am) run_march2<1, 2, true, false, false>(p, sums); else run_march2<1, 2, false, false, false>(p, sums); return 0; }
        if (C == 3 && S == 1) { if (am) run_march2<3, 1, true, false, false>(p, sums); else run_march2<3, 1, false, false, false>(p, sums); return 0; }
        if (C == 3 && S == 2) { if (am) run_march2<3, 2, true, false, false>(p, sums); else run_march2<3, 2, false, false, false>(p, sums); return 0; }
        return 1;
    }
    if (variant == 2 && BWD) {
        const bool am = p.automask != nullptr;
        if (p.dbg) {   // test hook: the instantiations that also export the discrete decisions (S = 2)
            if (C == 1 && S == 2) { if (am) run_march2<1, 2, true, true>(p, sums); else run_march2<1, 2, false, true>(p, sums); return 0; }
            if (C == 3 && S == 2) { if (am) run_march2<3, 2, true, true>(p, sums); else run_march2<3, 2, false, true>(p, sums); return 0; }
            return 1;
        }
        if (C == 1 && S == 1) { if (am) run_march2<1, 1, true>(p, sums); else run_march2<1, 1, false>(p, sums); return 0; }
        if (C == 1 && S == 2) { if (am) run_march2<1, 2, true>(p, sums); else run_march2<1, 2, false>(p, sums); return 0; }
        if (C == 3 && S == 1) { if (am) run_march2<3, 1, true>(p, sums); else run_march2<3, 1, false>(p, sums); return 0; }
        if (C == 3 && S == 2) { if (am) run_march2<3, 2, true>(p, sums); else run_march2<3, 2, false>(p, sums); return 0; }
        return 1;
    }
    return 1;
}

// mode: 0 fwd, 1 bwd (uses d->saved), 2 fwdbwd; variant: 0 tile phases, 1 marching warps (R rows per chunk)
static int emul_vsl(const md2_vsl_desc* d, int mode, float gloss, int variant, int R) {
    const int W = d->W, H = d->H, N = d->N, L = d->L, S = d->S, C = d->C;
    const bool bwd = mode != 0;
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.W = W; p.H = H; p.N = N; p.L = L; p.S = S;
    p.tgt = d->target; p.tgt_ns = d->target_image_stride;
    for (int s = 0; s < S; ++s) {
        p.src[s] = d->source[s]; p.src_ns[s] = d->source_image_stride[s];
        p.gsrc[s] = bwd ? d->grad_source[s] : nullptr;
        p.viz_warped[s] = d->viz_warped[s];
    }
    p.viz_loss = d->viz_loss; p.automask = d->automask;
    p.dbg = (bwd && variant == 2) ? d->debug_choices : nullptr;
    const float mind = (float)(1.0 / (double)d->max_depth), maxd = (float)(1.0 / (double)d->min_depth);
    p.depth_a = maxd - mind; p.depth_b = mind;
    for (int l = 0; l < L; ++l) {
        p.smooth_w[l] = d->smooth_weight[l];
        p.disp[l] = d->disparity[l]; p.dw[l] = d->disp_w[l]; p.dh[l] = d->disp_h[l];
        p.gdisp[l] = bwd ? d->grad_disparity[l] : nullptr;
    }
    // full-resolution views of every scale (what the prep kernel / ctx scratch provide on the device)
    std::vector<std::vector<float>> dscr(L), gscr(L);
    for (int l = 0; l < L; ++l) {
        const bool native = p.dw[l] == W && p.dh[l] == H;
        if (native) { p.dfull[l] = p.disp[l]; p.gfull[l] = p.gdisp[l]; continue; }
        dscr[l].resize((size_t)N * W * H);
        gscr[l].assign((size_t)N * W * H, NAN);
        const float usx = up_scale(p.dw[l], W), usy = up_scale(p.dh[l], H);
        for (int n = 0; n < N; ++n)
            for (int gy = 0; gy < H; ++gy)
                for (int gx = 0; gx < W; ++gx)
                    dscr[l][((size_t)n * H + gy) * W + gx] =
                        disp_fullres(p.disp[l] + (size_t)n * p.dw[l] * p.dh[l], p.dw[l], p.dh[l], false, W, H, gx, gy);
        p.dfull[l] = dscr[l].data();
        p.gfull[l] = bwd ? gscr[l].data() : nullptr;
    }
    p.loss_scale = d->loss_scale; p.gloss = gloss; p.normalize_disp = d->normalize_disparity;
    p.mode = mode;
    p.pose.mode = d->pose_mode; p.pose.K = d->K; p.pose.invK = d->invK;
    for (int s = 0; s < S; ++s) {
        p.pose.rot[s] = d->rot[s]; p.pose.trans[s] = d->trans[s]; p.pose.invert[s] = d->invert[s];
        p.pose.grot[s] = d->grad_rot[s]; p.pose.gtrans[s] = d->grad_trans[s];
    }
    std::vector<float> ab((size_t)24 * S * N);
    for (int s = 0; s < S; ++s)
        for (int n = 0; n < N; ++n) prepare_pose_one(p.pose, s, n, ab.data() + ((size_t)s * N + n) * 12, ab.data() + ((size_t)(S + s) * N + n) * 12);
    p.pose_ab = ab.data();
    p.pose_e = ab.data() + (size_t)12 * S * N;
    p.pose_slot = 0;
    p.m_R = R > 0 ? R : 32;

    std::vector<float> stats((size_t)L * N * NSTAT, 0.f);
    if (mode == 1) {
        memcpy(stats.data(), d->saved, sizeof(float) * stats.size());
    } else if (mode == 2) {
        for (int n = 0; n < N; ++n) {
            std::vector<float> v(3 * L, 0.f);
            for (int gy = 0; gy < H; ++gy)
                for (int gx = 0; gx < W; ++gx) {
                    if (C == 1) stats_pixel_all<1>(p, n, gx, gy, v.data());
                    else stats_pixel_all<3>(p, n, gx, gy, v.data());
                }
            for (int l = 0; l < L; ++l)
                for (int k = 0; k < 3; ++k) stats[((size_t)l * N + n) * NSTAT + 1 + k] = v[3 * l + k];
        }
    }
    p.stats = stats.data();
    std::vector<double> sums;
    const int NP = NSTAT + 12 * S;
    if (bwd && variant == 2 && (p.viz_loss || p.viz_warped[0] || p.viz_warped[1])) {
        // as run_vsl on the device: the visualisation outputs come from the forward-only kernel
        std::vector<double> tmp;
        if (dispatch<false>(C, S, p, tmp, 2)) return 1;
    }
    if (bwd ? dispatch<true>(C, S, p, sums, variant) : dispatch<false>(C, S, p, sums, variant)) return 1;
    if (bwd)   // adjoint of the upsample for the low-res scales (down_adjoint_kernel on the device)
        for (int l = 0; l < L; ++l) {
            if (p.dw[l] == W && p.dh[l] == H) continue;
            const int w = p.dw[l], h = p.dh[l];
            for (int n = 0; n < N; ++n)
                for (int i = 0; i < w * h; ++i)
                    p.gdisp[l][(size_t)n * w * h + i] = upsample_adjoint_at(p.gfull[l] + (size_t)n * W * H, w, h, W, H, i % w, i / w);
        }

    if (mode != 1) {
        for (int z = 0; z < L * N; ++z) {
            stats[(size_t)z * NSTAT] = (float)sums[(size_t)z * NP];
            if (mode == 0)
                for (int k = 1; k < 4; ++k) stats[(size_t)z * NSTAT + k] = (float)sums[(size_t)z * NP + k];
        }
        if (d->saved) memcpy(d->saved, stats.data(), sizeof(float) * stats.size());
        if (d->loss) *d->loss = loss_from_stats(stats.data(), W, H, N, L, d->smooth_weight, d->loss_scale, d->normalize_disparity);
    }
    if (bwd) {
        for (int s = 0; s < S; ++s)
            for (int n = 0; n < N; ++n) {
                double G[9] = {0}, h[3] = {0};
                for (int l = 0; l < L; ++l) {
                    const double* su = sums.data() + ((size_t)l * N + n) * NP + NSTAT + 12 * s;
                    for (int k = 0; k < 9; ++k) G[k] += su[k];
                    for (int k = 0; k < 3; ++k) h[k] += su[9 + k];
                }
                finalize_pose(p.pose, s, n, G, h);
            }
    }
    return 0;
}

extern "C" int md2_emul_vsl(const md2_vsl_desc* d, int mode, float gloss) { return emul_vsl(d, mode, gloss, 0, 0); }
extern "C" int md2_emul_march(const md2_vsl_desc* d, int mode, float gloss, int R) { return emul_vsl(d, mode, gloss, 1, R); }
extern "C" int md2_emul_march2(const md2_vsl_desc* d, int mode, float gloss, int R) { return emul_vsl(d, mode, gloss, 2, R); }
