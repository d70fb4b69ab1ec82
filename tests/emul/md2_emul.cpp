// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or shipped with the product.
// Runs the __host__ __device__ tile phases of monodepth2.jl_b200/csrc/md2_fused.cuh sequentially
// on the CPU (one "thread" at a time, phase by phase) so that the halo / reflect / fold / reduction
// logic of the fused CUDA kernel can be checked against the oracle in a container without a GPU.
// All pointers in the descriptor are HOST pointers here.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/md2.h"
#include "../../monodepth2.jl_b200/csrc/md2_fused.cuh"

using namespace md2;

static void cm3(const float* m, double* o) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) o[3 * i + j] = (double)m[3 * j + i];
}

static void upsample(const float* in, float* out, int w, int h, int W, int H, int CN) {
    const float sx = up_scale(w, W), sy = up_scale(h, H);
    for (long long i = 0; i < (long long)W * H * CN; ++i) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const long long cn = i / ((long long)W * H);
        int x0, x1, y0, y1; float fx, fy;
        up_taps(x, sx, w, x0, x1, fx);
        up_taps(y, sy, h, y0, y1, fy);
        const float* b = in + cn * w * h;
        out[i] = bilerp(b[y0 * w + x0], b[y0 * w + x1], b[y1 * w + x0], b[y1 * w + x1], fx, fy);
    }
}
static void upsample_bwd(const float* gout, float* gin, int w, int h, int W, int H, int CN) {
    const float sx = up_scale(w, W), sy = up_scale(h, H);
    memset(gin, 0, sizeof(float) * (size_t)w * h * CN);
    for (long long i = 0; i < (long long)W * H * CN; ++i) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const long long cn = i / ((long long)W * H);
        int x0, x1, y0, y1; float fx, fy;
        up_taps(x, sx, w, x0, x1, fx);
        up_taps(y, sy, h, y0, y1, fy);
        float* b = gin + cn * w * h;
        const float g = gout[i];
        b[y0 * w + x0] += (1 - fx) * (1 - fy) * g; b[y0 * w + x1] += fx * (1 - fy) * g;
        b[y1 * w + x0] += (1 - fx) * fy * g;       b[y1 * w + x1] += fx * fy * g;
    }
}

template <int C, int S, bool BWD>
static void run_tiles(const FusedParams& p, std::vector<float>& sums) {
    using F = Fused<C, S, BWD>;
    const int NP = F::NPART;
    const int tw = (p.W + TILE_W - 1) / TILE_W, th = (p.H + TILE_H - 1) / TILE_H;
    sums.assign((size_t)p.L * p.N * NP, 0.f);
    std::vector<float> sm(F::SMEM_FLOATS);
    std::vector<FusedAcc<S>> acc(FUSED_THREADS);
    for (int z = 0; z < p.L * p.N; ++z) {
        const int scale = z / p.N, n = z % p.N;
        for (int by = 0; by < th; ++by)
            for (int bx = 0; bx < tw; ++bx) {
                const int tx0 = bx * TILE_W, ty0 = by * TILE_H;
                for (auto& v : sm) v = NAN;   // poison: catches reads of never-written smem
                for (auto& a : acc) a.clear();
                for (int t = 0; t < FUSED_THREADS; ++t) F::phase_load(p, sm.data(), scale, n, tx0, ty0, t, FUSED_THREADS);
                for (int t = 0; t < FUSED_THREADS; ++t) F::phase_windows(p, sm.data(), scale, n, tx0, ty0, t, FUSED_THREADS, acc[t]);
                if (!BWD) {
                    for (int t = 0; t < FUSED_THREADS; ++t) F::phase_smooth_fwd(p, sm.data(), tx0, ty0, t, FUSED_THREADS, acc[t]);
                } else {
                    for (int t = 0; t < FUSED_THREADS; ++t) F::phase_pixel_bwd(p, sm.data(), scale, n, tx0, ty0, t, FUSED_THREADS, acc[t]);
                }
                float* su = sums.data() + (size_t)z * NP;
                for (int t = 0; t < FUSED_THREADS; ++t) {
                    su[0] += acc[t].warp_sum; su[1] += acc[t].sx; su[2] += acc[t].sy; su[3] += acc[t].dsum;
                    if (BWD)
                        for (int s = 0; s < S; ++s)
                            for (int k = 0; k < 12; ++k) su[NSTAT + 12 * s + k] += acc[t].pose[s][k];
                }
            }
    }
}

template <bool BWD>
static int dispatch(int C, int S, const FusedParams& p, std::vector<float>& sums) {
    if (C == 1 && S == 1) { run_tiles<1, 1, BWD>(p, sums); return 0; }
    if (C == 1 && S == 2) { run_tiles<1, 2, BWD>(p, sums); return 0; }
    if (C == 3 && S == 1) { run_tiles<3, 1, BWD>(p, sums); return 0; }
    if (C == 3 && S == 2) { run_tiles<3, 2, BWD>(p, sums); return 0; }
    return 1;
}

// mode: 0 fwd, 1 bwd (uses d->saved), 2 fwdbwd
extern "C" int md2_emul_vsl(const md2_vsl_desc* d, int mode, float gloss) {
    const int W = d->W, H = d->H, N = d->N, L = d->L, S = d->S, C = d->C;
    const long long HW = (long long)W * H;
    const bool bwd = mode != 0;
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.W = W; p.H = H; p.N = N; p.L = L;
    p.tgt = d->target; p.tgt_ns = d->target_image_stride;
    for (int s = 0; s < S; ++s) {
        p.src[s] = d->source[s]; p.src_ns[s] = d->source_image_stride[s];
        p.gsrc[s] = bwd ? d->grad_source[s] : nullptr;
        p.viz_warped[s] = d->viz_warped[s];
    }
    p.viz_loss = d->viz_loss; p.automask = d->automask;
    const float mind = (float)(1.0 / (double)d->max_depth), maxd = (float)(1.0 / (double)d->min_depth);
    p.depth_a = maxd - mind; p.depth_b = mind;
    for (int l = 0; l < L; ++l) p.smooth_w[l] = d->smooth_weight[l];
    p.loss_scale = d->loss_scale; p.gloss = gloss; p.normalize_disp = d->normalize_disparity;

    // poses
    std::vector<float> ab((size_t)12 * S * N);
    double K[9], Ki[9];
    cm3(d->K, K); cm3(d->invK, Ki);
    for (int s = 0; s < S; ++s)
        for (int n = 0; n < N; ++n) {
            double R[9], t[3];
            if (d->pose_mode == 0) {
                cm3(d->rot[s] + 9 * n, R);
                for (int k = 0; k < 3; ++k) t[k] = d->trans[s][3 * n + k];
            } else {
                double r[3], tv[3];
                for (int k = 0; k < 3; ++k) { r[k] = d->rot[s][3 * n + k]; tv[k] = d->trans[s][3 * n + k]; }
                compose_T(r, tv, d->invert[s], R, t);
            }
            precompose(K, Ki, R, t, ab.data() + ((size_t)s * N + n) * 12);
        }
    p.pose_ab = ab.data();

    std::vector<std::vector<float>> up(L), gup(L);
    for (int l = 0; l < L; ++l) {
        if (d->disp_w[l] != W || d->disp_h[l] != H) {
            up[l].resize((size_t)N * HW);
            upsample(d->disparity[l], up[l].data(), d->disp_w[l], d->disp_h[l], W, H, N);
            p.disp[l] = up[l].data();
            if (bwd) { gup[l].assign((size_t)N * HW, 0.f); p.gdisp[l] = gup[l].data(); }
        } else {
            p.disp[l] = d->disparity[l];
            p.gdisp[l] = bwd ? d->grad_disparity[l] : nullptr;
        }
    }
    std::vector<float> stats((size_t)L * N * NSTAT, 0.f);
    if (mode == 1) {
        memcpy(stats.data(), d->saved, sizeof(float) * stats.size());
    } else if (mode == 2) {
        for (int z = 0; z < L * N; ++z) {
            const int scale = z / N, n = z % N;
            float sx = 0, sy = 0, ds = 0;
            for (long long i = 0; i < HW; ++i) {
                if (C == 1) stats_pixel<1>(p.disp[scale] + n * HW, d->target + n * d->target_image_stride, i, W, H, sx, sy, ds);
                else stats_pixel<3>(p.disp[scale] + n * HW, d->target + n * d->target_image_stride, i, W, H, sx, sy, ds);
            }
            stats[(size_t)z * NSTAT + 1] = sx; stats[(size_t)z * NSTAT + 2] = sy; stats[(size_t)z * NSTAT + 3] = ds;
        }
    }
    p.stats = stats.data();
    std::vector<float> sums;
    const int NP = NSTAT + 12 * S;
    if (bwd ? dispatch<true>(C, S, p, sums) : dispatch<false>(C, S, p, sums)) return 1;

    if (mode != 1) {
        for (int z = 0; z < L * N; ++z) {
            stats[(size_t)z * NSTAT] = sums[(size_t)z * NP];
            if (mode == 0)
                for (int k = 1; k < 4; ++k) stats[(size_t)z * NSTAT + k] = sums[(size_t)z * NP + k];
        }
        if (d->saved) memcpy(d->saved, stats.data(), sizeof(float) * stats.size());
        if (d->loss) *d->loss = loss_from_stats(stats.data(), W, H, N, L, d->smooth_weight, d->loss_scale, d->normalize_disparity);
    }
    if (bwd) {
        for (int s = 0; s < S; ++s)
            for (int n = 0; n < N; ++n) {
                double G[9] = {0}, h[3] = {0};
                for (int l = 0; l < L; ++l) {
                    const float* su = sums.data() + ((size_t)l * N + n) * NP + NSTAT + 12 * s;
                    for (int k = 0; k < 9; ++k) G[k] += su[k];
                    for (int k = 0; k < 3; ++k) h[k] += su[9 + k];
                }
                double Rub[9], tub[3];
                precompose_bwd(K, Ki, G, h, Rub, tub);
                if (d->pose_mode == 0) {
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) d->grad_rot[s][9 * n + 3 * c + r] = (float)Rub[3 * r + c];
                    for (int k = 0; k < 3; ++k) d->grad_trans[s][3 * n + k] = (float)tub[k];
                } else {
                    double r[3], tv[3], rb[3], tb[3];
                    for (int k = 0; k < 3; ++k) { r[k] = d->rot[s][3 * n + k]; tv[k] = d->trans[s][3 * n + k]; }
                    compose_T_bwd(r, tv, d->invert[s], Rub, tub, rb, tb);
                    for (int k = 0; k < 3; ++k) { d->grad_rot[s][3 * n + k] = (float)rb[k]; d->grad_trans[s][3 * n + k] = (float)tb[k]; }
                }
            }
        for (int l = 0; l < L; ++l)
            if (d->disp_w[l] != W || d->disp_h[l] != H)
                upsample_bwd(gup[l].data(), d->grad_disparity[l], d->disp_w[l], d->disp_h[l], W, H, N);
    }
    return 0;
}
