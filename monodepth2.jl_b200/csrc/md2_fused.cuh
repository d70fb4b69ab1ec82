// Fused view-synthesis loss tile pipeline (the hot path):
//   disparity -> depth -> backproject -> pose -> project -> bilinear border warp of S source
//   frames -> SSIM(3x3, reflect pad)+L1 photometric -> per-pixel min over sources
//   [-> automask] -> mean, + edge-aware smoothness; forward and backward in one tile pass.
// Restates src/training.jl:42-70 (per-scale body of train_loss) and its Zygote pullback.
//
// The phases are plain __host__ __device__ functions operating on a "shared memory" float
// buffer, separated by barriers in the __global__ wrapper (md2_kernels.cu).  tests/emul runs
// the same phases sequentially on the CPU to check the tile/halo/reflect logic without a GPU.
#pragma once
#include "md2_math.cuh"

namespace md2 {

constexpr int MAX_S = 2;   // source frames (reference: source_ids = [1,3])
constexpr int MAX_L = 8;   // decoder scales (reference: 4)
constexpr int TILE_W = 32;
constexpr int TILE_H = 16;
constexpr int FUSED_THREADS = 256;
constexpr int NSTAT = 4;   // per (scale, image): warp sum, smooth-x sum, smooth-y sum, disparity sum

struct FusedParams {
    int W, H, N, L;
    // target frame (N,C,H,W) view: element (n,c,y,x) at tgt[n*tgt_ns + c*H*W + y*W + x]
    const float* tgt; long long tgt_ns;
    const float* src[MAX_S]; long long src_ns[MAX_S];
    float* gsrc[MAX_S];                 // nullable; same strides as src; accumulated (atomics)
    const float* disp[MAX_L];           // full-resolution disparity (N,H,W) per scale
    float* gdisp[MAX_L];                // (N,H,W) per scale, written
    const float* automask;              // (N,H,W) or null   (src/training.jl:60-62)
    const float* pose_ab;               // (S,N,12) pre-composed A|b
    const float* stats;                 // (L,N,NSTAT) forward sums, needed by the backward pass
    float* partial;                     // (blocks, NPART) per-block partial sums
    float depth_a, depth_b;             // z = 1/(a d + b)   (src/utils.jl:175-179)
    float smooth_w[MAX_L];              // disparity_smoothness * scale_i  (src/training.jl:66-67)
    float loss_scale;                   // 1/L  (src/training.jl:77)
    float gloss;                        // upstream cotangent of the scalar loss
    int normalize_disp;                 // 1: d / (mean d + 1e-7) before smoothness (src/training.jl:64-65)
    float* viz_warped[MAX_S];           // optional (N,C,H,W) contiguous, last scale only
    float* viz_loss;                    // optional (N,H,W), last scale only
};

template <int S>
struct FusedAcc {              // per-thread accumulators, block-reduced at the end
    float warp_sum, sx, sy, dsum;
    float pose[S][12];         // G (9, row-major) | h (3)
    MD2_HD void clear() {
        warp_sum = sx = sy = dsum = 0.f;
        for (int s = 0; s < S; ++s)
            for (int k = 0; k < 12; ++k) pose[s][k] = 0.f;
    }
};

#if defined(__CUDA_ARCH__)
#define MD2_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define MD2_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

template <int C, int S, bool BWD>
struct Fused {
    static constexpr int HALO = BWD ? 2 : 1;
    static constexpr int RW = TILE_W + 2 * HALO, RH = TILE_H + 2 * HALO;  // pixel region
    static constexpr int QW = TILE_W + 2 * (HALO - 1), QH = TILE_H + 2 * (HALO - 1);  // windows
    static constexpr int RN = RW * RH, QN = QW * QH;
    static constexpr int NPART = NSTAT + 12 * S;
    // shared-memory carve-up (floats)
    static constexpr int OFF_WARPED = 0;                  // [S*C][RN]
    static constexpr int OFF_TGT = OFF_WARPED + S * C * RN;  // [C][RN]
    static constexpr int OFF_DISP = OFF_TGT + C * RN;     // [RN]
    static constexpr int OFF_COEF = OFF_DISP + RN;        // BWD: [3*C][QN] of the selected source
    static constexpr int OFF_SEL = OFF_COEF + (BWD ? 3 * C * QN : 0);  // BWD: [QN] selected source or -1
    static constexpr int SMEM_FLOATS = OFF_SEL + (BWD ? QN : 0);

    // warp one pixel (image coords gx,gy 0-based) from source s; returns C values.
    // If DERIV, also returns d/dix, d/diy per channel and the taps/projection.
    template <bool DERIV>
    static MD2_HD void warp_pixel(const FusedParams& p, int n, int s, int gx, int gy, float d,
                                  const float* ab, float* val, float* dix, float* diy, Taps& tp,
                                  Proj& pr, float& z) {
        z = 1.0f / fmaf(d, p.depth_a, p.depth_b);
        project_ab(ab, (float)(gx + 1), (float)(gy + 1), z, pr);
        tp = border_taps(pr.u, pr.v, p.W, p.H);
        const long long HW = (long long)p.W * p.H;
        const float* base = p.src[s] + (long long)n * p.src_ns[s];
        const int o00 = tp.y0 * p.W + tp.x0, o01 = tp.y0 * p.W + tp.x1;
        const int o10 = tp.y1 * p.W + tp.x0, o11 = tp.y1 * p.W + tp.x1;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float* b = base + c * HW;
            const float v00 = b[o00], v01 = b[o01], v10 = b[o10], v11 = b[o11];
            val[c] = bilerp(v00, v01, v10, v11, tp.fx, tp.fy);
            if (DERIV) {
                dix[c] = fmaf(tp.fy, (v11 - v10) - (v01 - v00), v01 - v00);
                diy[c] = fmaf(tp.fx, (v11 - v01) - (v10 - v00), v10 - v00);
            }
        }
    }

    // ---- phase 1: fill the pixel region (tile + halo): target, disparity, warped sources ----
    static MD2_HD void phase_load(const FusedParams& p, float* sm, int scale, int n, int tx0,
                                  int ty0, int tid, int nthreads) {
        const long long HW = (long long)p.W * p.H;
        const float* tg = p.tgt + (long long)n * p.tgt_ns;
        const float* dp = p.disp[scale] + (long long)n * HW;
        for (int i = tid; i < RN; i += nthreads) {
            const int lx = i % RW, ly = i / RW;
            int gx = tx0 - HALO + lx, gy = ty0 - HALO + ly;
            // reflect-pad(1): only -1 and W (resp. H) are ever read by an in-image window
            if (gx == -1) gx = 1;
            if (gx == p.W) gx = p.W - 2;
            if (gy == -1) gy = 1;
            if (gy == p.H) gy = p.H - 2;
            const bool ok = gx >= 0 && gx < p.W && gy >= 0 && gy < p.H;
            float d = 0.f;
            if (ok) d = dp[gy * p.W + gx];
            sm[OFF_DISP + i] = d;
#pragma unroll
            for (int c = 0; c < C; ++c) sm[OFF_TGT + c * RN + i] = ok ? tg[c * HW + gy * p.W + gx] : 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                float val[C];
#pragma unroll
                for (int c = 0; c < C; ++c) val[c] = 0.f;
                if (ok) {
                    Taps tp; Proj pr; float z;
                    warp_pixel<false>(p, n, s, gx, gy, d, p.pose_ab + ((long long)s * p.N + n) * 12,
                                      val, nullptr, nullptr, tp, pr, z);
                }
#pragma unroll
                for (int c = 0; c < C; ++c) sm[OFF_WARPED + (s * C + c) * RN + i] = val[c];
            }
        }
    }

    // ---- phase 2: per 3x3 window: SSIM + L1 -> photometric, min over sources, automask ----
    static MD2_HD void phase_windows(const FusedParams& p, float* sm, int scale, int n, int tx0,
                                     int ty0, int tid, int nthreads, FusedAcc<S>& acc) {
        const float up_photo = p.gloss * p.loss_scale / ((float)p.W * (float)p.H * (float)p.N);
        for (int i = tid; i < QN; i += nthreads) {
            const int qx = i % QW, qy = i / QW;
            const int gx = tx0 - (HALO - 1) + qx, gy = ty0 - (HALO - 1) + qy;
            const int lx = qx + 1, ly = qy + 1;   // position in the pixel region
            const bool inside = gx >= 0 && gx < p.W && gy >= 0 && gy < p.H;
            int best = -1;
            if (inside) {
                const int ctr = ly * RW + lx;
                float ysum[C], yysum[C], yc[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float* T = sm + OFF_TGT + c * RN;
                    yc[c] = T[ctr];
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            const float y = T[ctr + dy * RW + dx] - yc[c];
                            a += y;
                            b = fmaf(y, y, b);
                        }
                    ysum[c] = a; yysum[c] = b;
                }
                float pe_best = 0.f;
                float cf[3 * C];   // coefficients of the best source so far
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    float ssum = 0.f, lsum = 0.f;
                    float cs[3 * C];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float* X = sm + OFF_WARPED + (s * C + c) * RN;
                        const float* T = sm + OFF_TGT + c * RN;
                        const float xc = X[ctr];
                        float sx = 0.f, sxx = 0.f, sxy = 0.f;
#pragma unroll
                        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int o = ctr + dy * RW + dx;
                                const float x = X[o] - xc;
                                const float y = T[o] - yc[c];
                                sx += x;
                                sxx = fmaf(x, x, sxx);
                                sxy = fmaf(x, y, sxy);
                            }
                        const SsimWin w = ssim_window<BWD>(xc, yc[c], sx, ysum[c], sxx, yysum[c], sxy);
                        ssum += w.s;
                        lsum += fabsf(yc[c] - xc);
                        if (BWD) {
                            cs[3 * c + 0] = w.alpha * w.pass;
                            cs[3 * c + 1] = w.beta * w.pass;
                            cs[3 * c + 2] = w.gamma * w.pass;
                        }
                    }
                    const float pe = PHOTO_ALPHA * (ssum * (1.0f / C)) + (1.0f - PHOTO_ALPHA) * (lsum * (1.0f / C));
                    if (s == 0 || pe < pe_best) {   // strict <: first index wins ties (findmin)
                        pe_best = pe; best = s;
                        if (BWD) {
#pragma unroll
                            for (int k = 0; k < 3 * C; ++k) cf[k] = cs[k];
                        }
                    }
                }
                float wl = pe_best;
                if (p.automask) {
                    const float am = p.automask[(long long)n * p.W * p.H + gy * p.W + gx];
                    if (am <= wl) { wl = am; best = -1; }   // mask is first in the cat: wins ties
                }
                const bool in_tile = (qx >= HALO - 1) && (qx < HALO - 1 + TILE_W) &&
                                     (qy >= HALO - 1) && (qy < HALO - 1 + TILE_H);
                if (in_tile) {
                    acc.warp_sum += wl;
                    if (scale == p.L - 1) {
                        const long long o = (long long)n * p.W * p.H + gy * p.W + gx;
                        if (p.viz_loss) p.viz_loss[o] = wl;
#pragma unroll
                        for (int s = 0; s < S; ++s)
                            if (p.viz_warped[s]) {
#pragma unroll
                                for (int c = 0; c < C; ++c)
                                    p.viz_warped[s][((long long)n * C + c) * p.W * p.H + gy * p.W + gx] =
                                        sm[OFF_WARPED + (s * C + c) * RN + ctr];
                            }
                    }
                }
                if (BWD) {
                    const float k = (best >= 0) ? up_photo * (PHOTO_ALPHA / C) * (-0.5f) : 0.f;
#pragma unroll
                    for (int j = 0; j < 3 * C; ++j) sm[OFF_COEF + j * QN + i] = cf[j] * k;
                }
            } else if (BWD) {
#pragma unroll
                for (int j = 0; j < 3 * C; ++j) sm[OFF_COEF + j * QN + i] = 0.f;
            }
            if (BWD) sm[OFF_SEL + i] = (float)best;
        }
    }

    // ---- forward-only smoothness sums over tile pixels (FWD kernel and the stats pre-pass) ----
    static MD2_HD void smooth_terms(const FusedParams& p, const float* sm, int gx, int gy, int r,
                                    float& tx, float& ty) {
        // r = index of the pixel in the region; right neighbour r+1, lower neighbour r+RW
        const float d = sm[OFF_DISP + r];
        tx = 0.f; ty = 0.f;
        if (gx + 1 < p.W) {
            float g = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r] - sm[OFF_TGT + c * RN + r + 1]);
            tx = fabsf(d - sm[OFF_DISP + r + 1]) * expf(-g * (1.0f / C));
        }
        if (gy + 1 < p.H) {
            float g = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r] - sm[OFF_TGT + c * RN + r + RW]);
            ty = fabsf(d - sm[OFF_DISP + r + RW]) * expf(-g * (1.0f / C));
        }
    }

    static MD2_HD void phase_smooth_fwd(const FusedParams& p, float* sm, int tx0, int ty0, int tid,
                                        int nthreads, FusedAcc<S>& acc) {
        for (int i = tid; i < TILE_W * TILE_H; i += nthreads) {
            const int px = i % TILE_W, py = i / TILE_W;
            const int gx = tx0 + px, gy = ty0 + py;
            if (gx >= p.W || gy >= p.H) continue;
            const int r = (py + HALO) * RW + px + HALO;
            float tx, ty;
            smooth_terms(p, sm, gx, gy, r, tx, ty);
            acc.sx += tx; acc.sy += ty; acc.dsum += sm[OFF_DISP + r];
        }
    }

    static MD2_HD float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

    // ---- phase 3 (BWD): per tile pixel: d loss / d warped -> sampler -> projection -> disparity,
    //      pose accumulators, source-image scatter; plus the smoothness gradient ----
    static MD2_HD void phase_pixel_bwd(const FusedParams& p, float* sm, int scale, int n, int tx0,
                                       int ty0, int tid, int nthreads, FusedAcc<S>& acc) {
        const long long HW = (long long)p.W * p.H;
        const float up_photo = p.gloss * p.loss_scale / ((float)p.W * (float)p.H * (float)p.N);
        // smoothness constants for this (scale, image)
        const float* st = p.stats + ((long long)scale * p.N + n) * NSTAT;
        const float cx = 1.0f / ((float)(p.W - 1) * (float)p.H * (float)p.N);
        const float cy = 1.0f / ((float)p.W * (float)(p.H - 1) * (float)p.N);
        const float up_s = p.gloss * p.loss_scale * p.smooth_w[scale];
        float sA = up_s, sB = 0.f;
        if (p.normalize_disp) {
            const float m = st[3] / (float)HW + 1e-7f;
            sA = up_s / m;
            sB = up_s * (cx * st[1] + cy * st[2]) / (m * m * (float)HW);
        }
        for (int i = tid; i < TILE_W * TILE_H; i += nthreads) {
            const int px = i % TILE_W, py = i / TILE_W;
            const int gx = tx0 + px, gy = ty0 + py;
            if (gx >= p.W || gy >= p.H) continue;
            const int r = (py + HALO) * RW + px + HALO;   // pixel region index
            const int q = (py + 1) * QW + px + 1;         // window region index
            // 3x3 adjoint of (reflect-pad o mean-pool): fold weights double the border windows
            float sa[S][C], sb[S][C], sg[S][C];
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
                for (int c = 0; c < C; ++c) sa[s][c] = sb[s][c] = sg[s][c] = 0.f;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const float wy = 1.0f + ((gy == 1 && dy == -1) ? 1.f : 0.f) + ((gy == p.H - 2 && dy == 1) ? 1.f : 0.f);
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const float wx = 1.0f + ((gx == 1 && dx == -1) ? 1.f : 0.f) + ((gx == p.W - 2 && dx == 1) ? 1.f : 0.f);
                    const float w = wx * wy;
                    const int o = q + dy * QW + dx;
                    const float sel = sm[OFF_SEL + o];
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const float ws = (sel == (float)s) ? w : 0.f;
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            sa[s][c] = fmaf(ws, sm[OFF_COEF + (3 * c + 0) * QN + o], sa[s][c]);
                            sb[s][c] = fmaf(ws, sm[OFF_COEF + (3 * c + 1) * QN + o], sb[s][c]);
                            sg[s][c] = fmaf(ws, sm[OFF_COEF + (3 * c + 2) * QN + o], sg[s][c]);
                        }
                    }
                }
            }
            const float selj = sm[OFF_SEL + q];
            const float d = sm[OFF_DISP + r];
            float dbar_z = 0.f;   // d loss / d depth
            float zz = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                float ibar[C];
                bool any = false;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float xj = sm[OFF_WARPED + (s * C + c) * RN + r];
                    const float yj = sm[OFF_TGT + c * RN + r];
                    float g = fmaf(xj, sb[s][c], fmaf(yj, sg[s][c], sa[s][c]));
                    if (selj == (float)s) g += up_photo * ((1.0f - PHOTO_ALPHA) / C) * sgn(xj - yj);
                    ibar[c] = g;
                    any = any || (g != 0.f);
                }
                if (!any) continue;
                float val[C], dix[C], diy[C];
                Taps tp; Proj pr; float z;
                warp_pixel<true>(p, n, s, gx, gy, d, p.pose_ab + ((long long)s * p.N + n) * 12, val,
                                 dix, diy, tp, pr, z);
                zz = z;
                float du = 0.f, dv = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    du = fmaf(ibar[c], dix[c], du);
                    dv = fmaf(ibar[c], diy[c], dv);
                }
                du *= tp.mx; dv *= tp.my;
                if (p.gsrc[s]) {
                    float* gb = p.gsrc[s] + (long long)n * p.src_ns[s];
                    const float w00 = (1.f - tp.fx) * (1.f - tp.fy), w01 = tp.fx * (1.f - tp.fy);
                    const float w10 = (1.f - tp.fx) * tp.fy, w11 = tp.fx * tp.fy;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        float* gc = gb + c * HW;
                        MD2_ATOMIC_ADD(gc + tp.y0 * p.W + tp.x0, w00 * ibar[c]);
                        if (tp.x0 + 1 < p.W) MD2_ATOMIC_ADD(gc + tp.y0 * p.W + tp.x0 + 1, w01 * ibar[c]);
                        if (tp.y0 + 1 < p.H) MD2_ATOMIC_ADD(gc + (tp.y0 + 1) * p.W + tp.x0, w10 * ibar[c]);
                        if (tp.x0 + 1 < p.W && tp.y0 + 1 < p.H)
                            MD2_ATOMIC_ADD(gc + (tp.y0 + 1) * p.W + tp.x0 + 1, w11 * ibar[c]);
                    }
                }
                float cb[3];
                project_ab_bwd(pr, du, dv, cb);
                dbar_z += cb[0] * pr.ap[0] + cb[1] * pr.ap[1] + cb[2] * pr.ap[2];
                const float zp[3] = {z * (float)(gx + 1), z * (float)(gy + 1), z};
#pragma unroll
                for (int a = 0; a < 3; ++a) {
#pragma unroll
                    for (int b = 0; b < 3; ++b) acc.pose[s][3 * a + b] = fmaf(cb[a], zp[b], acc.pose[s][3 * a + b]);
                    acc.pose[s][9 + a] += cb[a];
                }
            }
            // depth -> disparity:  dz/dd = -a z^2
            float gd = 0.f;
            if (dbar_z != 0.f) gd = -p.depth_a * zz * zz * dbar_z;
            // smoothness gradient (src/utils.jl:159-173 with the mean-normalisation of
            // src/training.jl:64-65 folded in):  A ghat_j - B
            float gh = 0.f;
            {
                const float* D = sm + OFF_DISP;
                if (gx + 1 < p.W) {
                    float g = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r] - sm[OFF_TGT + c * RN + r + 1]);
                    gh += cx * sgn(D[r] - D[r + 1]) * expf(-g * (1.0f / C));
                }
                if (gx > 0) {
                    float g = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r - 1] - sm[OFF_TGT + c * RN + r]);
                    gh -= cx * sgn(D[r - 1] - D[r]) * expf(-g * (1.0f / C));
                }
                if (gy + 1 < p.H) {
                    float g = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r] - sm[OFF_TGT + c * RN + r + RW]);
                    gh += cy * sgn(D[r] - D[r + RW]) * expf(-g * (1.0f / C));
                }
                if (gy > 0) {
                    float g = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r - RW] - sm[OFF_TGT + c * RN + r]);
                    gh -= cy * sgn(D[r - RW] - D[r]) * expf(-g * (1.0f / C));
                }
            }
            gd += sA * gh - sB;
            p.gdisp[scale][(long long)n * HW + gy * p.W + gx] = gd;
        }
    }
};

}  // namespace md2

namespace md2 {

// one pixel's contribution to the smoothness sums and the disparity sum (src/utils.jl:159-173)
template <int C>
MD2_HD void stats_pixel(const float* d, const float* t, long long i, int W, int H, float& sx,
                        float& sy, float& ds) {
    const long long HW = (long long)W * H;
    const int x = (int)(i % W), y = (int)(i / W);
    const float dc = d[i];
    ds += dc;
    if (x + 1 < W) {
        float g = 0.f;
        for (int c = 0; c < C; ++c) g += fabsf(t[c * HW + i] - t[c * HW + i + 1]);
        sx += fabsf(dc - d[i + 1]) * expf(-g * (1.0f / C));
    }
    if (y + 1 < H) {
        float g = 0.f;
        for (int c = 0; c < C; ++c) g += fabsf(t[c * HW + i] - t[c * HW + i + W]);
        sy += fabsf(dc - d[i + W]) * expf(-g * (1.0f / C));
    }
}

// loss = loss_scale * sum_i [ mean(warp_loss_i) + smooth_w_i * smooth_loss(dhat_i, target) ]
// from the per-(scale,image) sums {warp, Sx, Sy, dsum}   (src/training.jl:64-69,77)
MD2_HD float loss_from_stats(const float* stats, int W, int H, int N, int L, const float* smooth_w,
                             float loss_scale, int normalize_disp) {
    const double P = (double)W * H;
    const double cx = 1.0 / ((double)(W - 1) * H * N), cy = 1.0 / ((double)W * (H - 1) * N);
    double loss = 0.0;
    for (int i = 0; i < L; ++i) {
        double warp = 0.0, sm = 0.0;
        for (int n = 0; n < N; ++n) {
            const float* st = stats + ((long long)i * N + n) * NSTAT;
            warp += st[0];
            const double m = normalize_disp ? ((double)st[3] / P + 1e-7) : 1.0;
            sm += (cx * st[1] + cy * st[2]) / m;
        }
        loss += warp / (P * N) + (double)smooth_w[i] * sm;
    }
    return (float)(loss * loss_scale);
}

}  // namespace md2
