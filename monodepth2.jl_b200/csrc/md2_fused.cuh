// Fused view-synthesis loss tile pipeline (the hot path), v2:
//   low-res disparity -> (align-corners bilinear upsample, A17) -> depth -> backproject -> pose ->
//   project -> bilinear border warp of S source frames -> SSIM(3x3, reflect pad)+L1 photometric
//   -> per-pixel min over sources [-> automask] -> mean, + edge-aware smoothness;
//   forward and backward in ONE tile pass, all decoder scales in one launch.
// Restates src/training.jl:42-70 (per-scale body of train_loss) and its Zygote pullback.
//
// A block of 128 threads owns a TW x TH tile.  Phases (separated by block barriers):
//   1 load     every pixel of tile+halo: target, disparity, both warped sources -> shared memory
//              (+ d warped / d coordinate of the tile pixels, so the backward never re-gathers)
//   2 windows  one thread per window column, marching down rows with rolling separable 3x3 sums:
//              SSIM, photometric error, arg-min over sources, automask -> loss partial sums and
//              (backward) per-window SSIM gradient coefficients
//   3 pixels   (backward) one thread per pixel column, marching down rows: rolling 3x3 adjoint
//              sums of the coefficients, sampler / projection / depth adjoints, pose
//              accumulators, source-image scatter with vertically carried + warp-merged atomics,
//              smoothness gradient
//   4 downsample (backward, low-res scales) adjoint of the bilinear upsample: separable gather
//              of the tile's full-res disparity gradient into the low-res patch it touches
// The phases are plain __host__ __device__ functions over a "shared memory" float buffer;
// tests/emul runs them sequentially on the CPU so the tile / halo / reflect / rolling logic is
// checked against the CPU reference without a GPU (warp-shuffle-only optimisations are
// device-only and covered by the GPU parity tests).
#pragma once
#include "md2_math.cuh"

namespace md2 {

constexpr int MAX_S = 2;   // source frames (reference: source_ids = [1,3])
constexpr int MAX_L = 8;   // decoder scales (reference: 4)
constexpr int FUSED_THREADS = 128;
constexpr int NSTAT = 4;   // per (scale, image): warp sum, smooth-x sum, smooth-y sum, disparity sum

struct PoseIO {            // pose inputs / gradient outputs, Julia memory order
    int mode;              // 0: rot = R (3,3,N) col-major, trans = t ; 1: rot = rvec (3,N), trans = tvec
    const float* rot[MAX_S];
    const float* trans[MAX_S];
    int invert[MAX_S];
    const float* K; const float* invK;   // (3,3) col-major
    float* grot[MAX_S];
    float* gtrans[MAX_S];
};

struct FusedParams {
    int W, H, N, L, S;
    // target frame (N,C,H,W) view: element (n,c,y,x) at tgt[n*tgt_ns + c*H*W + y*W + x]
    const float* tgt; long long tgt_ns;
    const float* src[MAX_S]; long long src_ns[MAX_S];
    float* gsrc[MAX_S];                 // nullable; same strides as src; accumulated (atomics)
    const float* disp[MAX_L];           // disparity (N,dh,dw) per scale at its NATIVE size
    int dw[MAX_L], dh[MAX_L];
    float* gdisp[MAX_L];                // (N,dh,dw): written (full-res) / accumulated (low-res, pre-zeroed)
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
    // device-side finalisation by the last block (md2_fused.cu)
    int mode;                           // 0 fwd, 1 bwd, 2 fwdbwd
    unsigned int* counters;             // [L*N + 1], zero on entry, left zero on exit
    float* sums;                        // (L*N, NPART)
    float* stats_out;                   // (L*N, NSTAT) written in modes 0 and 2
    float* saved;                       // nullable copy of stats_out for a later bwd
    float* loss;                        // nullable device scalar
    PoseIO pose;
    // marching-warp kernel (md2_march.cuh): rows per chunk, offset of this call's pose rows in the
    // constant-memory pose table
    int m_R, pose_slot;
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
__device__ __forceinline__ float md2_rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define MD2_RCP_FAST(x) md2_rcp_approx(x)
#define MD2_EXP(x) __expf(x)
#else
#define MD2_ATOMIC_ADD(p, v) (*(p) += (v))
#define MD2_RCP_FAST(x) (1.0f / (x))
#define MD2_EXP(x) expf(x)
#endif

// reciprocal for the geometry (depth, perspective divide): approx + one Newton step, <= 1 ulp
MD2_HD float rcp_acc(float x) {
#if defined(__CUDA_ARCH__)
    const float r = md2_rcp_approx(x);
    return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
    return 1.0f / x;
#endif
}

MD2_HD float sgnf(float v) { return (float)(v > 0.f) - (float)(v < 0.f); }

// disparity of scale `l` at full-resolution pixel (gx,gy): direct read or on-the-fly
// align-corners bilinear upsample (A17)
MD2_HD float disp_fullres(const float* __restrict__ dp, int dw, int dh, bool native, float usx, float usy,
                          int W, int gx, int gy) {
    if (native) return dp[gy * W + gx];
    int x0, x1, y0, y1; float fx, fy;
    up_taps(gx, usx, dw, x0, x1, fx);
    up_taps(gy, usy, dh, y0, y1, fy);
    return bilerp(dp[y0 * dw + x0], dp[y0 * dw + x1], dp[y1 * dw + x0], dp[y1 * dw + x1], fx, fy);
}

// SSIM of one window from sums centred on (xr, yr), fast reciprocals; see ssim_window
template <bool WITH_COEF>
MD2_HD SsimWin ssim_window_fast(float xr, float yr, float sx, float sy, float sxx, float syy, float sxy) {
    const float r9 = 1.0f / 9.0f;
    const float dx = sx * r9, dy = sy * r9;
    const float mux = xr + dx, muy = yr + dy;
    const float vx = fmaf(-dx, dx, sxx * r9);
    const float vy = fmaf(-dy, dy, syy * r9);
    const float vxy = fmaf(-dx, dy, sxy * r9);
    const float A = 2.0f * mux * muy + SSIM_C1;
    const float B = 2.0f * vxy + SSIM_C2;
    const float Cc = (mux * mux + muy * muy) + SSIM_C1;
    const float D = (vx + vy) + SSIM_C2;
    const float rC = MD2_RCP_FAST(Cc), rD = MD2_RCP_FAST(D);
    const float inv = rC * rD;
    const float S = A * B * inv;
    const float raw = (1.0f - S) * 0.5f;
    SsimWin o;
    o.s = fminf(fmaxf(raw, 0.0f), 1.0f);
    if (WITH_COEF) {
        const float k = 2.0f / 9.0f;
        o.pass = (raw >= 0.0f && raw <= 1.0f) ? 1.0f : 0.0f;
        o.beta = -k * S * rD;
        o.gamma = k * A * inv;
        o.alpha = k * (muy * (B - A) * inv - S * mux * rC + S * mux * rD);
    } else {
        o.pass = 0.f; o.alpha = o.beta = o.gamma = 0.f;
    }
    return o;
}

template <int C, int S, bool BWD>
struct Fused {
    static constexpr int HALO = BWD ? 2 : 1;
    static constexpr int TW = BWD ? 30 : 32;          // BWD: 30 + 2 window halo = 32 = one warp of window columns
    static constexpr int TH = 16;
    static constexpr int RW = TW + 2 * HALO, RH = TH + 2 * HALO;          // pixel region
    static constexpr int QW = TW + 2 * (HALO - 1), QH = TH + 2 * (HALO - 1);  // window region
    static constexpr int RN = RW * RH, QN = QW * QH, TN = TW * TH;
    static constexpr int NPART = NSTAT + 12 * S;
    static constexpr int NSTRIP = FUSED_THREADS / 32;   // row strips in phases 2 and 3
    static_assert(QW == 32, "one warp per window row");
    static constexpr int PATCH_MAX = TW / 2 + 3;        // low-res patch width bound (scale <= 1/2)
    // shared-memory carve-up (floats)
    static constexpr int OFF_WARPED = 0;                        // [S*C][RN]
    static constexpr int OFF_TGT = OFF_WARPED + S * C * RN;     // [C][RN]
    static constexpr int OFF_DISP = OFF_TGT + C * RN;           // [RN]
    static constexpr int OFF_COEF = OFF_DISP + RN;              // BWD: [3*C][QN] of the selected source
    static constexpr int OFF_SEL = OFF_COEF + (BWD ? 3 * C * QN : 0);   // BWD: [QN] selected source or -1
    static constexpr int OFF_SLOPE = OFF_SEL + (BWD ? QN : 0);  // BWD: [S*C*2][TN] d warped / d (ix, iy)
    static constexpr int OFF_GD = OFF_SLOPE + (BWD ? 2 * S * C * TN : 0);   // BWD: [TN] full-res disparity gradient
    static constexpr int OFF_TMP = OFF_GD + (BWD ? TN : 0);     // BWD: [NSEG][TH][PATCH_MAX] phase-4 scratch
    static constexpr int OFF_TAPX = OFF_TMP + (BWD ? 4 * TH * PATCH_MAX : 0);   // BWD: [TW] x0 (as float), [TW] fx
    static constexpr int OFF_TAPY = OFF_TAPX + (BWD ? 2 * TW : 0);          // BWD: [TH] y0, [TH] fy
    static constexpr int SMEM_FLOATS = OFF_TAPY + (BWD ? 2 * TH : 0);

    struct Warped {
        float val[C], dix[C], diy[C];
    };

    // geometry of one pixel for source s (pose row `ab`)
    static MD2_HD void project_pixel(const FusedParams& p, const float* ab, int gx, int gy,
                                     float z, Proj& pr, Taps& tp) {
        const float px = (float)(gx + 1), py = (float)(gy + 1);
        pr.ap[0] = fmaf(ab[0], px, fmaf(ab[1], py, ab[2]));
        pr.ap[1] = fmaf(ab[3], px, fmaf(ab[4], py, ab[5]));
        pr.ap[2] = fmaf(ab[6], px, fmaf(ab[7], py, ab[8]));
        const float c0 = fmaf(z, pr.ap[0], ab[9]);
        const float c1 = fmaf(z, pr.ap[1], ab[10]);
        const float c2 = fmaf(z, pr.ap[2], ab[11]);
        pr.q = rcp_acc(c2 + PROJ_EPS);
        pr.u = c0 * pr.q;
        pr.v = c1 * pr.q;
        tp = border_taps(pr.u, pr.v, p.W, p.H);
    }

    template <bool DERIV>
    static MD2_HD void gather(const FusedParams& p, int n, int s, const Taps& tp, Warped& w) {
        const long long HW = (long long)p.W * p.H;
        const float* base = p.src[s] + (long long)n * p.src_ns[s];
        const float* r0 = base + (tp.y0 * p.W + tp.x0);   // the 2x2 cell is always inside the image
        const float* r1 = r0 + p.W;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float v00 = r0[c * HW], v01 = r0[c * HW + 1], v10 = r1[c * HW], v11 = r1[c * HW + 1];
            w.val[c] = bilerp(v00, v01, v10, v11, tp.fx, tp.fy);
            if (DERIV) {
                w.dix[c] = fmaf(tp.fy, (v11 - v10) - (v01 - v00), v01 - v00);
                w.diy[c] = fmaf(tp.fx, (v11 - v01) - (v10 - v00), v10 - v00);
            }
        }
    }

    // image coordinates of region pixel i (reflect-pad(1): only -1 and W, resp. H, are ever read
    // by an in-image window); false if the pixel is outside the image
    static MD2_HD bool region_coords(const FusedParams& p, int i, int tx0, int ty0, int& gx, int& gy) {
        const int lx = i % RW, ly = i / RW;
        gx = tx0 - HALO + lx; gy = ty0 - HALO + ly;
        if (gx == -1) gx = 1;
        if (gx == p.W) gx = p.W - 2;
        if (gy == -1) gy = 1;
        if (gy == p.H) gy = p.H - 2;
        return gx >= 0 && gx < p.W && gy >= 0 && gy < p.H;
    }

    // ---- phase 1 ----
    static MD2_HD void phase_load(const FusedParams& p, float* sm, int scale, int n, int tx0, int ty0, int tid) {
        const long long HW = (long long)p.W * p.H;
        const float* tg = p.tgt + (long long)n * p.tgt_ns;
        const int dw = p.dw[scale], dh = p.dh[scale];
        const bool native = (dw == p.W && dh == p.H);
        const float* dp = p.disp[scale] + (long long)n * dw * dh;
        const float usx = up_scale(dw, p.W), usy = up_scale(dh, p.H);
        float ab[S][12];   // pose rows in registers for the whole loop
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int k = 0; k < 12; ++k) ab[s][k] = p.pose_ab[((long long)s * p.N + n) * 12 + k];
        // pass A: target + disparity of every region pixel (independent loads, all in flight)
#pragma unroll
        for (int k = 0; k < (RN + FUSED_THREADS - 1) / FUSED_THREADS; ++k) {
            const int i = tid + k * FUSED_THREADS;
            if (i < RN) {
                int gx, gy;
                const bool ok = region_coords(p, i, tx0, ty0, gx, gy);
                sm[OFF_DISP + i] = ok ? disp_fullres(dp, dw, dh, native, usx, usy, p.W, gx, gy) : 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) sm[OFF_TGT + c * RN + i] = ok ? tg[c * HW + gy * p.W + gx] : 0.f;
            }
        }
        // pass B: geometry + gather of both sources (each thread re-reads its own disparities)
#pragma unroll 2
        for (int i = tid; i < RN; i += FUSED_THREADS) {
            const int lx = i % RW, ly = i / RW;
            int gx, gy;
            const bool ok = region_coords(p, i, tx0, ty0, gx, gy);
            const float d = sm[OFF_DISP + i];
            const float z = rcp_acc(fmaf(d, p.depth_a, p.depth_b));
            const bool in_tile = BWD && lx >= HALO && lx < HALO + TW && ly >= HALO && ly < HALO + TH;
            const int ti = (ly - HALO) * TW + (lx - HALO);
#pragma unroll
            for (int s = 0; s < S; ++s) {
                Warped w;
#pragma unroll
                for (int c = 0; c < C; ++c) { w.val[c] = 0.f; w.dix[c] = 0.f; w.diy[c] = 0.f; }
                if (ok) {
                    Taps tp; Proj pr;
                    project_pixel(p, ab[s], gx, gy, z, pr, tp);
                    gather<BWD>(p, n, s, tp, w);
                }
#pragma unroll
                for (int c = 0; c < C; ++c) sm[OFF_WARPED + (s * C + c) * RN + i] = w.val[c];
                if (BWD && in_tile) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        sm[OFF_SLOPE + ((s * C + c) * 2 + 0) * TN + ti] = w.dix[c];
                        sm[OFF_SLOPE + ((s * C + c) * 2 + 1) * TN + ti] = w.diy[c];
                    }
                }
            }
        }
        if (BWD && !native) {   // upsample taps of the tile's columns / rows for phase 4
            for (int i = tid; i < TW + TH; i += FUSED_THREADS) {
                int a0, a1; float f;
                if (i < TW) {
                    const int gx = tx0 + i < p.W ? tx0 + i : p.W - 1;
                    up_taps(gx, usx, dw, a0, a1, f);
                    sm[OFF_TAPX + i] = (float)a0;
                    sm[OFF_TAPX + TW + i] = (a1 > a0) ? f : 0.f;   // clamped last column: all weight on a0
                } else {
                    const int j = i - TW;
                    const int gy = ty0 + j < p.H ? ty0 + j : p.H - 1;
                    up_taps(gy, usy, dh, a0, a1, f);
                    sm[OFF_TAPY + j] = (float)a0;
                    sm[OFF_TAPY + TH + j] = (a1 > a0) ? f : 0.f;
                }
            }
        }
    }

    // horizontal 3-sums of one region row for the window column starting at region column `c0`
    struct RowSums {
        float hx[S][C], hxx[S][C], hxy[S][C], hy[C], hyy[C];
        float xm[S][C], ym[C];   // the (un-centred) middle values: the window centre one row later
    };

    static MD2_HD void row_sums(const float* sm, int row, int c0, const float (&xr)[S][C], const float (&yr)[C],
                                RowSums& o) {
        const int b = row * RW + c0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float* T = sm + OFF_TGT + c * RN + b;
            const float t1 = T[1];
            const float y0 = T[0] - yr[c], y1 = t1 - yr[c], y2 = T[2] - yr[c];
            o.ym[c] = t1;
            o.hy[c] = y0 + y1 + y2;
            o.hyy[c] = fmaf(y2, y2, fmaf(y1, y1, y0 * y0));
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float* X = sm + OFF_WARPED + (s * C + c) * RN + b;
                const float v1 = X[1];
                const float x0 = X[0] - xr[s][c], x1 = v1 - xr[s][c], x2 = X[2] - xr[s][c];
                o.xm[s][c] = v1;
                o.hx[s][c] = x0 + x1 + x2;
                o.hxx[s][c] = fmaf(x2, x2, fmaf(x1, x1, x0 * x0));
                o.hxy[s][c] = fmaf(x2, y2, fmaf(x1, y1, x0 * y0));
            }
        }
    }

    // ---- phase 2: one thread per window column, rolling over its strip of window rows ----
    static MD2_HD void phase_windows(const FusedParams& p, float* sm, int scale, int n, int tx0, int ty0, int tid,
                                     FusedAcc<S>& acc) {
        const int col = tid & 31, strip = tid >> 5;
        const int r0 = (strip * QH) / NSTRIP, r1 = ((strip + 1) * QH) / NSTRIP;   // window rows [r0, r1)
        const int gx = tx0 - (HALO - 1) + col;
        const bool col_in = gx >= 0 && gx < p.W;
        const float up_photo = p.gloss * p.loss_scale / ((float)p.W * (float)p.H * (float)p.N);
        const bool col_tile = (col >= HALO - 1) && (col < HALO - 1 + TW);
        // centring reference: the strip's first window centre (any constant works; a local value
        // keeps the centred squares small)
        float xr[S][C], yr[C];
        {
            const int ctr0 = (r0 + 1) * RW + col + 1;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                yr[c] = sm[OFF_TGT + c * RN + ctr0];
#pragma unroll
                for (int s = 0; s < S; ++s) xr[s][c] = sm[OFF_WARPED + (s * C + c) * RN + ctr0];
            }
        }
        RowSums a, b, cur;   // region rows rr-2, rr-1, rr
        row_sums(sm, r0, col, xr, yr, a);
        row_sums(sm, r0 + 1, col, xr, yr, b);
        for (int q = r0; q < r1; ++q) {          // window row q uses region rows q, q+1, q+2
            row_sums(sm, q + 2, col, xr, yr, cur);
            const int gy = ty0 - (HALO - 1) + q;
            const int qi = q * QW + col;
            const bool inside = col_in && gy >= 0 && gy < p.H;
            int best = -1;
            float cf[3 * C];
#pragma unroll
            for (int k = 0; k < 3 * C; ++k) cf[k] = 0.f;
            if (inside) {
                float pe_best = 0.f;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    float ssum = 0.f, lsum = 0.f;
                    float cs[3 * C];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const SsimWin w = ssim_window_fast<BWD>(
                            xr[s][c], yr[c], a.hx[s][c] + b.hx[s][c] + cur.hx[s][c], a.hy[c] + b.hy[c] + cur.hy[c],
                            a.hxx[s][c] + b.hxx[s][c] + cur.hxx[s][c], a.hyy[c] + b.hyy[c] + cur.hyy[c],
                            a.hxy[s][c] + b.hxy[s][c] + cur.hxy[s][c]);
                        ssum += w.s;
                        lsum += fabsf(b.ym[c] - b.xm[s][c]);
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
                const bool in_tile = col_tile && (q >= HALO - 1) && (q < HALO - 1 + TH);
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
                                    p.viz_warped[s][((long long)n * C + c) * p.W * p.H + gy * p.W + gx] = b.xm[s][c];
                            }
                    }
                }
            }
            if (BWD) {
                const float k = (best >= 0) ? up_photo * (PHOTO_ALPHA / C) * (-0.5f) : 0.f;
#pragma unroll
                for (int j = 0; j < 3 * C; ++j) sm[OFF_COEF + j * QN + qi] = cf[j] * k;
                sm[OFF_SEL + qi] = (float)best;
            }
            a = b; b = cur;
        }
    }

    // ---- forward-only smoothness sums over the tile pixels ----
    static MD2_HD void phase_smooth_fwd(const FusedParams& p, float* sm, int tx0, int ty0, int tid, FusedAcc<S>& acc) {
        for (int i = tid; i < TN; i += FUSED_THREADS) {
            const int px = i % TW, py = i / TW;
            const int gx = tx0 + px, gy = ty0 + py;
            if (gx >= p.W || gy >= p.H) continue;
            const int r = (py + HALO) * RW + px + HALO;
            const float d = sm[OFF_DISP + r];
            if (gx + 1 < p.W) {
                float g = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r] - sm[OFF_TGT + c * RN + r + 1]);
                acc.sx += fabsf(d - sm[OFF_DISP + r + 1]) * MD2_EXP(-g * (1.0f / C));
            }
            if (gy + 1 < p.H) {
                float g = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r] - sm[OFF_TGT + c * RN + r + RW]);
                acc.sy += fabsf(d - sm[OFF_DISP + r + RW]) * MD2_EXP(-g * (1.0f / C));
            }
            acc.dsum += d;
        }
    }

    // weighted horizontal 3-sums of the coefficient maps of one window row (adjoint of
    // reflect-pad o mean-pool: wl / wr double the image-border windows)
    struct CoefRow {
        float t[3 * C];    // all windows
        float s0[3 * C];   // windows whose selected source is 0
    };
    static MD2_HD void coef_row(const float* sm, int qrow, int px, float wl, float wr, CoefRow& o) {
        const int b = qrow * QW + px;
        const float e0 = sm[OFF_SEL + b], e1 = sm[OFF_SEL + b + 1], e2 = sm[OFF_SEL + b + 2];
        const float m0 = (e0 == 0.f) ? wl : 0.f, m1 = (e1 == 0.f) ? 1.f : 0.f, m2 = (e2 == 0.f) ? wr : 0.f;
#pragma unroll
        for (int j = 0; j < 3 * C; ++j) {
            const float* cm = sm + OFF_COEF + j * QN + b;
            const float c0 = cm[0], c1 = cm[1], c2 = cm[2];
            o.t[j] = fmaf(wl, c0, fmaf(wr, c2, c1));
            if (S > 1) o.s0[j] = fmaf(m0, c0, fmaf(m2, c2, m1 * c1));
            else o.s0[j] = 0.f;
        }
    }

    // emit the source-image gradient of one tap pair row (x0,y) / (x0+1,y)
    static MD2_HD void red2(const FusedParams& p, float* gb, int x0, int y, const float (&v0)[C], const float (&v1)[C],
                            bool emit1) {
        const int HW = p.W * p.H;
        float* a = gb + (y * p.W + x0);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            MD2_ATOMIC_ADD(a + c * HW, v0[c]);
            if (emit1) MD2_ATOMIC_ADD(a + c * HW + 1, v1[c]);
        }
    }

    // ---- phase 3 (BWD): one thread per pixel column, rolling over its strip of tile rows ----
    static MD2_HD void phase_pixel_bwd(const FusedParams& p, float* sm, int scale, int n, int tx0, int ty0, int tid,
                                       FusedAcc<S>& acc) {
        const long long HW = (long long)p.W * p.H;
        const int px = tid & 31, strip = tid >> 5;
        const int py0 = (strip * TH) / NSTRIP, py1 = ((strip + 1) * TH) / NSTRIP;
        const int gx = tx0 + px;
        const bool col_ok = px < TW && gx < p.W;
        const int pxc = px < TW ? px : TW - 1;   // idle lanes read valid shared memory
        const float up_photo = p.gloss * p.loss_scale / ((float)p.W * (float)p.H * (float)p.N);
        const bool native = (p.dw[scale] == p.W && p.dh[scale] == p.H);
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
        const float wl = (gx == 1) ? 2.f : 1.f, wr = (gx == p.W - 2) ? 2.f : 1.f;
        float ab[S][12];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int k = 0; k < 12; ++k) ab[s][k] = p.pose_ab[((long long)s * p.N + n) * 12 + k];

        // vertical carry of the lower tap pair of the previous row, per source
        float car0[S][C], car1[S][C];
        int cx0[S], cy0[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            cx0[s] = -1; cy0[s] = -1;
#pragma unroll
            for (int c = 0; c < C; ++c) { car0[s][c] = 0.f; car1[s][c] = 0.f; }
        }

        CoefRow ra, rb, rc;   // window rows py, py+1, py+2 (window-region coordinates)
        coef_row(sm, py0, pxc, wl, wr, ra);
        coef_row(sm, py0 + 1, pxc, wl, wr, rb);
        for (int py = py0; py < py1; ++py) {
            coef_row(sm, py + 2, pxc, wl, wr, rc);
            const int gy = ty0 + py;
            const bool valid = col_ok && gy < p.H;
            const float wu = (gy == 1) ? 2.f : 1.f, wd = (gy == p.H - 2) ? 2.f : 1.f;
            const int r = (py + HALO) * RW + pxc + HALO;   // pixel region index
            const int ti = py * TW + pxc;
            const float selj = sm[OFF_SEL + (py + 1) * QW + pxc + 1];
            const float d = sm[OFF_DISP + r];
            const float z = rcp_acc(fmaf(d, p.depth_a, p.depth_b));
            float dbar_z = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                // d loss / d warped_s at this pixel
                float ibar[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    float sa, sb, sg;
                    const float ta = fmaf(wu, ra.t[3 * c], fmaf(wd, rc.t[3 * c], rb.t[3 * c]));
                    const float tb = fmaf(wu, ra.t[3 * c + 1], fmaf(wd, rc.t[3 * c + 1], rb.t[3 * c + 1]));
                    const float tg = fmaf(wu, ra.t[3 * c + 2], fmaf(wd, rc.t[3 * c + 2], rb.t[3 * c + 2]));
                    if (S == 1) { sa = ta; sb = tb; sg = tg; }
                    else {
                        const float za = fmaf(wu, ra.s0[3 * c], fmaf(wd, rc.s0[3 * c], rb.s0[3 * c]));
                        const float zb = fmaf(wu, ra.s0[3 * c + 1], fmaf(wd, rc.s0[3 * c + 1], rb.s0[3 * c + 1]));
                        const float zg = fmaf(wu, ra.s0[3 * c + 2], fmaf(wd, rc.s0[3 * c + 2], rb.s0[3 * c + 2]));
                        if (s == 0) { sa = za; sb = zb; sg = zg; }
                        else { sa = ta - za; sb = tb - zb; sg = tg - zg; }
                    }
                    const float xj = sm[OFF_WARPED + (s * C + c) * RN + r];
                    const float yj = sm[OFF_TGT + c * RN + r];
                    float g = fmaf(xj, sb, fmaf(yj, sg, sa));
                    if (selj == (float)s) g += up_photo * ((1.0f - PHOTO_ALPHA) / C) * sgnf(xj - yj);
                    ibar[c] = valid ? g : 0.f;
                }
                bool act = false;
#pragma unroll
                for (int c = 0; c < C; ++c) act = act || (ibar[c] != 0.f);
                Taps tp;
                tp.x0 = 0; tp.y0 = 0; tp.x1 = 0; tp.y1 = 0; tp.fx = 0.f; tp.fy = 0.f; tp.mx = 0.f; tp.my = 0.f;
                if (act) {   // (sources that were not selected anywhere in the 3x3 neighbourhood skip all of this)
                    Proj pr;
                    project_pixel(p, ab[s], gx, gy, z, pr, tp);
                    float du = 0.f, dv = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        du = fmaf(ibar[c], sm[OFF_SLOPE + ((s * C + c) * 2 + 0) * TN + ti], du);
                        dv = fmaf(ibar[c], sm[OFF_SLOPE + ((s * C + c) * 2 + 1) * TN + ti], dv);
                    }
                    du *= tp.mx; dv *= tp.my;
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
                const bool sval = valid && act;   // this pixel scatters into source s
                // source-image gradient: scatter with vertical carry (+ warp merge on the device)
                if (p.gsrc[s]) {
                    float* gb = p.gsrc[s] + (long long)n * p.src_ns[s];
                    const float w00 = (1.f - tp.fx) * (1.f - tp.fy), w01 = tp.fx * (1.f - tp.fy);
                    const float w10 = (1.f - tp.fx) * tp.fy, w11 = tp.fx * tp.fy;
                    float t0[C], t1[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) { t0[c] = w00 * ibar[c]; t1[c] = w01 * ibar[c]; }
                    const bool have = cx0[s] >= 0;
                    const bool aligned = have && sval && tp.x0 == cx0[s] && tp.y0 == cy0[s] + 1;
                    if (aligned) {
#pragma unroll
                        for (int c = 0; c < C; ++c) { t0[c] += car0[s][c]; t1[c] += car1[s][c]; }
                    } else if (have) {
                        red2(p, gb, cx0[s], cy0[s] + 1, car0[s], car1[s], true);
                    }
                    bool emit1 = sval;
#if defined(__CUDA_ARCH__)
                    {   // merge with the horizontal neighbours: my right tap is the right lane's left tap
                        const int key = sval ? ((tp.y0 << 16) | tp.x0) : -2;
                        const int key_r = __shfl_down_sync(0xffffffffu, key, 1);
                        const int key_l = __shfl_up_sync(0xffffffffu, key, 1);
                        const int lane = threadIdx.x & 31;
                        const bool absorbed = sval && lane < 31 && key_r == key + 1;
                        const bool absorb = sval && lane > 0 && key_l >= 0 && key_l + 1 == key;
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const float fl = __shfl_up_sync(0xffffffffu, t1[c], 1);
                            if (absorb) t0[c] += fl;
                        }
                        if (absorbed) emit1 = false;
                    }
#endif
                    if (sval) red2(p, gb, tp.x0, tp.y0, t0, t1, emit1);
                    if (sval) {
                        cx0[s] = tp.x0; cy0[s] = tp.y0;
#pragma unroll
                        for (int c = 0; c < C; ++c) { car0[s][c] = w10 * ibar[c]; car1[s][c] = w11 * ibar[c]; }
                    } else {
                        cx0[s] = -1;
                    }
                }
            }
            // depth -> disparity:  dz/dd = -a z^2
            float gd = -p.depth_a * z * z * dbar_z;
            // smoothness gradient (src/utils.jl:159-173 with the mean-normalisation of
            // src/training.jl:64-65 folded in):  A ghat_j - B
            float gh = 0.f;
            {
                const float* D = sm + OFF_DISP;
                if (gx + 1 < p.W) {
                    float g = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r] - sm[OFF_TGT + c * RN + r + 1]);
                    gh += cx * sgnf(D[r] - D[r + 1]) * MD2_EXP(-g * (1.0f / C));
                }
                if (gx > 0) {
                    float g = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r - 1] - sm[OFF_TGT + c * RN + r]);
                    gh -= cx * sgnf(D[r - 1] - D[r]) * MD2_EXP(-g * (1.0f / C));
                }
                if (gy + 1 < p.H) {
                    float g = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r] - sm[OFF_TGT + c * RN + r + RW]);
                    gh += cy * sgnf(D[r] - D[r + RW]) * MD2_EXP(-g * (1.0f / C));
                }
                if (gy > 0) {
                    float g = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) g += fabsf(sm[OFF_TGT + c * RN + r - RW] - sm[OFF_TGT + c * RN + r]);
                    gh -= cy * sgnf(D[r - RW] - D[r]) * MD2_EXP(-g * (1.0f / C));
                }
            }
            gd += sA * gh - sB;
            if (!valid) gd = 0.f;
            if (native) {
                if (valid) p.gdisp[scale][(long long)n * HW + gy * p.W + gx] = gd;
            } else if (px < TW) {
                sm[OFF_GD + ti] = gd;
            }
            ra = rb; rb = rc;
        }
        // flush the carried lower tap pairs of the strip's last row
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (p.gsrc[s] && cx0[s] >= 0)
                red2(p, p.gsrc[s] + (long long)n * p.src_ns[s], cx0[s], cy0[s] + 1, car0[s], car1[s], true);
    }

    // ---- phase 4 (BWD, low-res scale): adjoint of the bilinear upsample, separable gather ----
    // 4a: tmp[y][ex] = sum_x wx(x, ex) gd[y][x]      4b: out[ey][ex] += sum_y wy(y, ey) tmp[y][ex]
    static MD2_HD void patch_extent(const float* sm, int& ex0, int& nex, int& ey0, int& ney) {
        ex0 = (int)sm[OFF_TAPX];
        nex = (int)sm[OFF_TAPX + TW - 1] + 2 - ex0;   // up to the last column's x0 + 1
        ey0 = (int)sm[OFF_TAPY];
        ney = (int)sm[OFF_TAPY + TH - 1] + 2 - ey0;
        if (nex > PATCH_MAX) nex = PATCH_MAX;          // cannot happen for decoder scales <= 1/2
    }
    // 4a: thread (row y, segment g): horizontal pass over its NSEG-th of the tile columns into its own
    //     slice tmp[g][y][.]; 4b: thread (patch column e, segment g of the rows): vertical pass over the
    //     sum of the slices, flushed with one atomic per finished low-res element
    static constexpr int NSEG = 4;
    static MD2_HD void phase_down_a(const FusedParams& p, float* sm, int tid) {
        int ex0, nex, ey0, ney;
        patch_extent(sm, ex0, nex, ey0, ney);
        constexpr int SEGW = (TW + NSEG - 1) / NSEG;
        for (int i = tid; i < TH * NSEG; i += FUSED_THREADS) {
            const int y = i / NSEG, g = i % NSEG;
            float* row = sm + OFF_TMP + (g * TH + y) * PATCH_MAX;
            const int xa = g * SEGW, xb = (xa + SEGW < TW) ? xa + SEGW : TW;
            for (int e = 0; e < nex; ++e) row[e] = 0.f;
            int cur = (int)sm[OFF_TAPX + xa] - ex0;   // patch column of accumulator a0 (a1 is cur + 1)
            float a0 = 0.f, a1 = 0.f;
            for (int x = xa; x < xb; ++x) {
                const int e = (int)sm[OFF_TAPX + x] - ex0;
                const float f = sm[OFF_TAPX + TW + x], gdv = sm[OFF_GD + y * TW + x];
                while (cur < e) {        // columns only advance: flush the finished one
                    if (cur < PATCH_MAX) row[cur] = a0;
                    a0 = a1; a1 = 0.f; ++cur;
                }
                a0 = fmaf(1.f - f, gdv, a0);
                a1 = fmaf(f, gdv, a1);
            }
            if (cur < PATCH_MAX) row[cur] = a0;
            if (cur + 1 < PATCH_MAX) row[cur + 1] = a1;
        }
    }
    static MD2_HD void phase_down_b(const FusedParams& p, float* sm, int scale, int n, int tid) {
        int ex0, nex, ey0, ney;
        patch_extent(sm, ex0, nex, ey0, ney);
        const int dw = p.dw[scale], dh = p.dh[scale];
        float* g = p.gdisp[scale] + (long long)n * dw * dh;
        constexpr int SEGH = (TH + NSEG - 1) / NSEG;
        for (int i = tid; i < nex * NSEG; i += FUSED_THREADS) {
            const int e = i / NSEG, sg = i % NSEG;
            if (ex0 + e >= dw) continue;
            const int ya = sg * SEGH, yb = (ya + SEGH < TH) ? ya + SEGH : TH;
            int cur = (int)sm[OFF_TAPY + ya];
            float a0 = 0.f, a1 = 0.f;   // accumulators of low-res rows cur and cur+1
            for (int y = ya; y < yb; ++y) {
                const int r = (int)sm[OFF_TAPY + y];
                const float f = sm[OFF_TAPY + TH + y];
                float v = 0.f;
#pragma unroll
                for (int k = 0; k < NSEG; ++k) v += sm[OFF_TMP + (k * TH + y) * PATCH_MAX + e];
                while (cur < r) {   // rows only advance; flush the finished one
                    if (a0 != 0.f && cur < dh) MD2_ATOMIC_ADD(g + cur * dw + ex0 + e, a0);
                    a0 = a1; a1 = 0.f; ++cur;
                }
                a0 = fmaf(1.f - f, v, a0);
                a1 = fmaf(f, v, a1);
            }
            if (a0 != 0.f && cur < dh) MD2_ATOMIC_ADD(g + cur * dw + ex0 + e, a0);
            if (a1 != 0.f && cur + 1 < dh) MD2_ATOMIC_ADD(g + (cur + 1) * dw + ex0 + e, a1);
        }
    }
};

// one full-resolution pixel's contribution to the smoothness sums and the disparity sum of
// every scale (src/utils.jl:159-173): the edge weights exp(-|dT|) are shared by all scales
template <int C>
MD2_HD void stats_pixel_all(const FusedParams& p, int n, int gx, int gy, float* v /*[L][3]: sx, sy, dsum*/) {
    const long long HW = (long long)p.W * p.H;
    const float* t = p.tgt + (long long)n * p.tgt_ns + gy * p.W + gx;
    const bool hx = gx + 1 < p.W, hy = gy + 1 < p.H;
    float wx = 0.f, wy = 0.f;
    if (hx) {
        float g = 0.f;
        for (int c = 0; c < C; ++c) g += fabsf(t[c * HW] - t[c * HW + 1]);
        wx = MD2_EXP(-g * (1.0f / C));
    }
    if (hy) {
        float g = 0.f;
        for (int c = 0; c < C; ++c) g += fabsf(t[c * HW] - t[c * HW + p.W]);
        wy = MD2_EXP(-g * (1.0f / C));
    }
    for (int l = 0; l < p.L; ++l) {
        const int dw = p.dw[l], dh = p.dh[l];
        const bool native = (dw == p.W && dh == p.H);
        const float* dp = p.disp[l] + (long long)n * dw * dh;
        const float usx = up_scale(dw, p.W), usy = up_scale(dh, p.H);
        const float d = disp_fullres(dp, dw, dh, native, usx, usy, p.W, gx, gy);
        v[3 * l + 2] += d;
        if (hx) v[3 * l + 0] += fabsf(d - disp_fullres(dp, dw, dh, native, usx, usy, p.W, gx + 1, gy)) * wx;
        if (hy) v[3 * l + 1] += fabsf(d - disp_fullres(dp, dw, dh, native, usx, usy, p.W, gx, gy + 1)) * wy;
    }
}

// one pixel's contribution to the smoothness sums and the disparity sum (stand-alone smooth_loss)
template <int C>
MD2_HD void stats_pixel(const float* d, const float* t, long long i, int W, int H, float& sx, float& sy, float& ds) {
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

// pose gradients of one (source, image) from the accumulated G|h of all scales
MD2_HD void finalize_pose(const PoseIO& io, int s, int n, const double* G, const double* h) {
    double K[9], Ki[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { K[3 * i + j] = (double)io.K[3 * j + i]; Ki[3 * i + j] = (double)io.invK[3 * j + i]; }
    double Rub[9], tub[3];
    precompose_bwd(K, Ki, G, h, Rub, tub);
    if (io.mode == 0) {
        if (io.grot[s])
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) io.grot[s][9 * n + 3 * c + r] = (float)Rub[3 * r + c];
        if (io.gtrans[s])
            for (int k = 0; k < 3; ++k) io.gtrans[s][3 * n + k] = (float)tub[k];
    } else {
        double r[3], tv[3], rb[3], tb[3];
        for (int k = 0; k < 3; ++k) { r[k] = io.rot[s][3 * n + k]; tv[k] = io.trans[s][3 * n + k]; }
        compose_T_bwd(r, tv, io.invert[s], Rub, tub, rb, tb);
        if (io.grot[s])
            for (int k = 0; k < 3; ++k) io.grot[s][3 * n + k] = (float)rb[k];
        if (io.gtrans[s])
            for (int k = 0; k < 3; ++k) io.gtrans[s][3 * n + k] = (float)tb[k];
    }
}

// [composeT] + pre-composition of one (source, image)
MD2_HD void prepare_pose_one(const PoseIO& io, int s, int n, float* ab) {
    double K[9], Ki[9], R[9], t[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { K[3 * i + j] = (double)io.K[3 * j + i]; Ki[3 * i + j] = (double)io.invK[3 * j + i]; }
    if (io.mode == 0) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[3 * i + j] = (double)io.rot[s][9 * n + 3 * j + i];
        for (int k = 0; k < 3; ++k) t[k] = io.trans[s][3 * n + k];
    } else {
        double r[3], tv[3];
        for (int k = 0; k < 3; ++k) { r[k] = io.rot[s][3 * n + k]; tv[k] = io.trans[s][3 * n + k]; }
        compose_T(r, tv, io.invert[s], R, t);
    }
    precompose(K, Ki, R, t, ab);
}

}  // namespace md2
