// Fused view-synthesis loss, "marching warp" pipeline (the hot path, v4).
//
// Same maths as md2_fused.cuh (src/training.jl:42-70 and its Zygote pullback), different mapping:
// ONE WARP owns a 32-column strip of one (scale, image) and marches down a chunk of rows; lane =
// image column.  Per row the warp runs three pipelined stages, all state in registers:
//   L(i)    disparity [bilinear upsample, A17] -> depth -> backproject/pose/project -> 4-tap border
//           gather of the S source frames; horizontal 3-sums for the SSIM windows come from the
//           neighbouring lanes by warp shuffle
//   W(i-1)  vertical rolling 3-sums -> SSIM + L1 photometric error, arg-min over sources, automask,
//           loss partial sums; (backward) the per-window SSIM gradient coefficients and their
//           horizontal adjoint 3-sums (shuffles)
//   P(i-2)  (backward) vertical adjoint sums -> d loss / d warped, sampler / projection / depth
//           adjoints, pose accumulators, source-image scatter (lower tap pair carried to the next
//           row, right tap merged into the right-hand lane), smoothness gradient, disparity
//           gradient (low-res scales: adjoint of the upsample, vertical in registers, horizontal
//           through a 64-float per-warp scratch)
// Nothing is shared between warps: no block barriers, the halo is 2 columns each side (28 of 32
// lanes produce outputs) and 2 rows at the chunk ends.  The only shared memory is a per-lane
// 3-row ring holding what P(i-2) needs from L(i-2) (sampler slopes, projected coordinates).
//
// The code is written against a tiny warp interface (w_up / w_dn / w_shfl / w_ballot / w_sync) so
// that tests/emul can run the very same source on the CPU with 32 cooperative fibers per warp
// (MD2_WARP_EMU, test infrastructure only).
#pragma once
#include <string.h>

#include "md2_fused.cuh"

namespace md2 {

struct alignas(16) Vec4 { float x, y, z, w; };

#if defined(MD2_WARP_EMU)
#define MD2_DEV inline
// emu_xchg / emu_ballot: tests/emul/warp_emu.h, included before this file
inline float w_shfl(float v, int src, int) { unsigned int u; memcpy(&u, &v, 4); u = emu_xchg(u, src); float r; memcpy(&r, &u, 4); return r; }
inline int w_shfl(int v, int src, int) { return (int)emu_xchg((unsigned int)v, src); }
inline float w_up(float v, int lane) { return w_shfl(v, lane > 0 ? lane - 1 : 0, lane); }
inline float w_dn(float v, int lane) { return w_shfl(v, lane < 31 ? lane + 1 : 31, lane); }
inline int w_up(int v, int lane) { return w_shfl(v, lane > 0 ? lane - 1 : 0, lane); }
inline int w_dn(int v, int lane) { return w_shfl(v, lane < 31 ? lane + 1 : 31, lane); }
inline unsigned int w_ballot(bool p) { return emu_ballot(p ? 1 : 0); }
inline void w_sync() { emu_ballot(0); }
inline int w_popc(unsigned int m) { return __builtin_popcount(m); }
#define MD2_POSE(p, idx) ((p).pose_ab[(idx)])
#else
#define MD2_DEV __device__ __forceinline__
// pre-composed pose rows (A | b per source and image), read through the constant cache so that the
// warp-uniform values live in uniform registers / constant operands instead of 24 vector registers
constexpr int POSE_CONST_FLOATS = 12288;   // 48 KB: S*N <= 1024 per call
constexpr int POSE_SLOT_FLOATS = 3072;     // per-ctx slot (S*N <= 256); larger calls use the whole table
__constant__ float c_pose[POSE_CONST_FLOATS];
#define MD2_POSE(p, idx) (c_pose[(idx)])
MD2_DEV float w_shfl(float v, int src, int) { return __shfl_sync(0xffffffffu, v, src); }
MD2_DEV int w_shfl(int v, int src, int) { return __shfl_sync(0xffffffffu, v, src); }
MD2_DEV float w_up(float v, int) { return __shfl_up_sync(0xffffffffu, v, 1); }
MD2_DEV float w_dn(float v, int) { return __shfl_down_sync(0xffffffffu, v, 1); }
MD2_DEV int w_up(int v, int) { return __shfl_up_sync(0xffffffffu, v, 1); }
MD2_DEV int w_dn(int v, int) { return __shfl_down_sync(0xffffffffu, v, 1); }
MD2_DEV unsigned int w_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
MD2_DEV void w_sync() { __syncwarp(); }
MD2_DEV int w_popc(unsigned int m) { return __popc(m); }
#endif

template <int C, int S, bool BWD>
struct March {
    static constexpr int HALO = BWD ? 2 : 1;
    static constexpr int OW = 32 - 2 * HALO;             // output columns per strip
    static constexpr int NPART = NSTAT + 12 * S;
    static constexpr int NSL4 = (2 * S * C + 3) / 4;     // Vec4 units of sampler slopes per pixel
    static constexpr int RING4 = NSL4 + S;               // + (u, v, q, z) per source
    static constexpr int SMEM_FLOATS = BWD ? (3 * RING4 * 32 * 4 + 64) : 4;   // ring + down-sampling scratch
    static_assert(NPART <= 32, "one lane per partial sum");

    struct RowSums {       // horizontal 3-sums of one pixel row, window column centred on this lane
        float hx[S][C], hxx[S][C], hxy[S][C], hy[C], hyy[C];
        float xm[S][C], ym[C];   // this lane's own (centred) values
    };
    struct CoefRow {       // horizontal adjoint 3-sums of the coefficient maps of one window row
        float t[3 * C];    // all windows
        float s0[3 * C];   // windows whose selected source is 0
    };

    // one warp, one (strip sx, chunk cy, scale*N+n = z) work item; on return lane-local partial
    // sums are in v[0..NPART)
    static MD2_DEV void run(const FusedParams& p, int sx, int cy, int z, int lane, float* wsm, float (&v)[32]) {
        const int W = p.W, H = p.H, HW = W * H;
        const int scale = z / p.N, n = z - scale * p.N;
        const int X0 = sx * OW, Y0 = cy * p.m_R;
        const int Y1 = (Y0 + p.m_R < H) ? Y0 + p.m_R : H;
        // this lane's column (reflect-pad(1): only -1 and W are ever used by an in-image window)
        const int gxr = X0 - HALO + lane;
        int gxm = gxr == -1 ? 1 : (gxr == W ? W - 2 : gxr);
        gxm = gxm < 0 ? 0 : (gxm > W - 1 ? W - 1 : gxm);
        const bool col_img = gxr >= 0 && gxr < W;
        const bool wcol = col_img && lane >= 1 && lane <= 30;               // window column
        const bool pcol = col_img && lane >= HALO && lane < 32 - HALO;      // output pixel column
        const float px = (float)(gxm + 1);

        const float* tgn = p.tgt + (long long)n * p.tgt_ns;
        const float* tg = tgn + gxm;
        const float* sb[S];
#pragma unroll
        for (int s = 0; s < S; ++s) sb[s] = p.src[s] + (long long)n * p.src_ns[s];
        const int dw = p.dw[scale], dh = p.dh[scale];
        const bool native = (dw == W && dh == H);
        const float* dp = p.disp[scale] + (long long)n * dw * dh;
        const float usx = up_scale(dw, W), usy = up_scale(dh, H);
        int xa0 = 0, xa1 = 0;
        float fxu = 0.f;
        if (!native) up_taps(gxm, usx, dw, xa0, xa1, fxu);

        // centring constant of the window sums (any constant is exact; a local value keeps the
        // centred squares small): the target at the middle of the strip chunk
        float rc[C];
        {
            const int ym = (Y0 + Y1) >> 1;
            const int xm = X0 + OW / 2 < W ? X0 + OW / 2 : W - 1;
#pragma unroll
            for (int c = 0; c < C; ++c) rc[c] = tgn[c * HW + ym * W + xm];
        }
        int pb[S];
        float apx[S][3];   // A[:,0] px + A[:,2]: the lane-constant part of A p
#pragma unroll
        for (int s = 0; s < S; ++s) {
            pb[s] = p.pose_slot + (s * p.N + n) * 12;
#pragma unroll
            for (int k = 0; k < 3; ++k) apx[s][k] = fmaf(MD2_POSE(p, pb[s] + 3 * k), px, MD2_POSE(p, pb[s] + 3 * k + 2));
        }

        const float up_photo = p.gloss * p.loss_scale / ((float)W * (float)H * (float)p.N);
        // backward-only constants
        const float cxn = 1.0f / ((float)(W - 1) * (float)H * (float)p.N);
        const float cyn = 1.0f / ((float)W * (float)(H - 1) * (float)p.N);
        float sA = 0.f, sB = 0.f;
        if (BWD) {
            const float* st = p.stats + ((long long)scale * p.N + n) * NSTAT;
            const float up_s = p.gloss * p.loss_scale * p.smooth_w[scale];
            sA = up_s;
            if (p.normalize_disp) {
                const float m = st[3] / (float)HW + 1e-7f;
                sA = up_s / m;
                sB = up_s * (cxn * st[1] + cyn * st[2]) / (m * m * (float)HW);
            }
        }
        const float wl = (gxr == 1) ? 2.f : 1.f, wr = (gxr == W - 2) ? 2.f : 1.f;
        const bool has_right = col_img && gxr + 1 < W;

        // adjoint of the upsample, horizontal part: lanes [pm,p0) feed low-res column b0+lane with
        // their right-tap weight, lanes [p0,pp) with their left-tap weight
        int b0 = 0, pm = 0, p0 = 0, pp = 0;
        float fxa = 0.f;
        if (BWD && !native) {
            const int xb0 = pcol ? xa0 : (lane < HALO ? -(1 << 20) : (1 << 20));
            fxa = (pcol && xa1 > xa0) ? fxu : 0.f;     // clamped last column: all weight on xa0
            b0 = w_shfl(xb0, HALO, lane);
            for (int k = -1; k <= 20; ++k) {
                const int cnt = w_popc(w_ballot(xb0 < b0 + k));
                if (k == lane - 1) pm = cnt;
                if (k == lane) p0 = cnt;
                if (k == lane + 1) pp = cnt;
            }
        }

        Vec4* ring = reinterpret_cast<Vec4*>(wsm);
        float* dsm = wsm + 3 * RING4 * 32 * 4;

        RowSums a, b, cur;
        CoefRow ra, rb, rcf;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            a.hy[c] = a.hyy[c] = a.ym[c] = b.hy[c] = b.hyy[c] = b.ym[c] = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                a.hx[s][c] = a.hxx[s][c] = a.hxy[s][c] = a.xm[s][c] = 0.f;
                b.hx[s][c] = b.hxx[s][c] = b.hxy[s][c] = b.xm[s][c] = 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < 3 * C; ++j) ra.t[j] = ra.s0[j] = rb.t[j] = rb.s0[j] = rcf.t[j] = rcf.s0[j] = 0.f;
        float Da = 0.f, Db = 0.f, ey_prev = 0.f;
        int sel_prev = -1;
        float warp_sum = 0.f, ssx = 0.f, ssy = 0.f, dsum = 0.f;
        float P0[S][3], P1[S][3], Ph[S][3];
        float car0[S][C], car1[S][C];
        int ckey[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            ckey[s] = -1;
#pragma unroll
            for (int k = 0; k < 3; ++k) P0[s][k] = P1[s][k] = Ph[s][k] = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) car0[s][c] = car1[s][c] = 0.f;
        }
        float da0 = 0.f, da1 = 0.f;
        int dcur = -1;
        int slot = 0;

        for (int i = Y0 - HALO; i < Y1 + HALO; ++i) {
            // ================= L(i): this lane's pixel of row i =================
            int gym = i == -1 ? 1 : (i == H ? H - 2 : i);
            gym = gym < 0 ? 0 : (gym > H - 1 ? H - 1 : gym);
            const float py = (float)(gym + 1);
            float d;
            if (native) {
                d = dp[gym * W + gxm];
            } else {
                int ya0, ya1; float fyu;
                up_taps(gym, usy, dh, ya0, ya1, fyu);
                d = bilerp(dp[ya0 * dw + xa0], dp[ya0 * dw + xa1], dp[ya1 * dw + xa0], dp[ya1 * dw + xa1], fxu, fyu);
            }
            const float zv = rcp_acc(fmaf(d, p.depth_a, p.depth_b));
            float Tc[C], Xc[S][C];
#pragma unroll
            for (int c = 0; c < C; ++c) Tc[c] = tg[c * HW + gym * W] - rc[c];
            float slopes[NSL4 * 4];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float ap0 = fmaf(MD2_POSE(p, pb[s] + 1), py, apx[s][0]);
                const float ap1 = fmaf(MD2_POSE(p, pb[s] + 4), py, apx[s][1]);
                const float ap2 = fmaf(MD2_POSE(p, pb[s] + 7), py, apx[s][2]);
                const float c0 = fmaf(zv, ap0, MD2_POSE(p, pb[s] + 9));
                const float c1 = fmaf(zv, ap1, MD2_POSE(p, pb[s] + 10));
                const float c2 = fmaf(zv, ap2, MD2_POSE(p, pb[s] + 11));
                const float q = rcp_acc(c2 + PROJ_EPS);
                const float u = c0 * q, vv = c1 * q;
                const Taps tp = border_taps(u, vv, W, H);
                const float* r0 = sb[s] + (tp.y0 * W + tp.x0);   // the 2x2 cell is always inside the image
                const float* r1 = r0 + W;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float v00 = r0[c * HW], v01 = r0[c * HW + 1], v10 = r1[c * HW], v11 = r1[c * HW + 1];
                    Xc[s][c] = bilerp(v00, v01, v10, v11, tp.fx, tp.fy) - rc[c];
                    if (BWD) {
                        slopes[(s * C + c) * 2 + 0] = fmaf(tp.fy, (v11 - v10) - (v01 - v00), v01 - v00);
                        slopes[(s * C + c) * 2 + 1] = fmaf(tp.fx, (v11 - v01) - (v10 - v00), v10 - v00);
                    }
                }
                if (BWD) {
                    Vec4 g4; g4.x = u; g4.y = vv; g4.z = q; g4.w = zv;
                    ring[(slot * RING4 + NSL4 + s) * 32 + lane] = g4;
                }
            }
            if (BWD) {
#pragma unroll
                for (int k = 2 * S * C; k < NSL4 * 4; ++k) slopes[k] = 0.f;
#pragma unroll
                for (int k = 0; k < NSL4; ++k) {
                    Vec4 s4; s4.x = slopes[4 * k]; s4.y = slopes[4 * k + 1]; s4.z = slopes[4 * k + 2]; s4.w = slopes[4 * k + 3];
                    ring[(slot * RING4 + k) * 32 + lane] = s4;
                }
            }
            // horizontal 3-sums (window column centred on this lane)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float yl = w_up(Tc[c], lane), yr = w_dn(Tc[c], lane);
                cur.ym[c] = Tc[c];
                cur.hy[c] = yl + Tc[c] + yr;
                cur.hyy[c] = fmaf(yr, yr, fmaf(Tc[c], Tc[c], yl * yl));
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const float xl = w_up(Xc[s][c], lane), xr = w_dn(Xc[s][c], lane);
                    cur.xm[s][c] = Xc[s][c];
                    cur.hx[s][c] = xl + Xc[s][c] + xr;
                    cur.hxx[s][c] = fmaf(xr, xr, fmaf(Xc[s][c], Xc[s][c], xl * xl));
                    cur.hxy[s][c] = fmaf(xr, yr, fmaf(Xc[s][c], Tc[c], xl * yl));
                }
            }

            // ================= W(i-1): windows centred on row i-1 =================
            int sel_q = -1;
            if (i >= Y0 - HALO + 2) {
                const int q = i - 1;
                const bool inside = wcol && q >= 0 && q < H;
                float cf[3 * C];
#pragma unroll
                for (int k = 0; k < 3 * C; ++k) cf[k] = 0.f;
                float wlv = 0.f;
                if (inside) {
                    float pe_best = 0.f;
                    float sy3[C], syy3[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        sy3[c] = a.hy[c] + b.hy[c] + cur.hy[c];
                        syy3[c] = a.hyy[c] + b.hyy[c] + cur.hyy[c];
                    }
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        float ssum = 0.f, lsum = 0.f;
                        float cs[3 * C];
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const SsimWin w = ssim_window_fast<BWD>(
                                rc[c], rc[c], a.hx[s][c] + b.hx[s][c] + cur.hx[s][c], sy3[c],
                                a.hxx[s][c] + b.hxx[s][c] + cur.hxx[s][c], syy3[c],
                                a.hxy[s][c] + b.hxy[s][c] + cur.hxy[s][c]);
                            ssum += w.s;
                            lsum += fabsf(b.ym[c] - b.xm[s][c]);
                            if (BWD) {   // coefficients for CENTRED member values: alpha' = alpha + r (beta + gamma)
                                cs[3 * c + 0] = fmaf(rc[c], w.beta + w.gamma, w.alpha) * w.pass;
                                cs[3 * c + 1] = w.beta * w.pass;
                                cs[3 * c + 2] = w.gamma * w.pass;
                            }
                        }
                        const float pe = PHOTO_ALPHA * (ssum * (1.0f / C)) + (1.0f - PHOTO_ALPHA) * (lsum * (1.0f / C));
                        if (s == 0 || pe < pe_best) {   // strict <: first index wins ties (findmin)
                            pe_best = pe; sel_q = s;
                            if (BWD) {
#pragma unroll
                                for (int k = 0; k < 3 * C; ++k) cf[k] = cs[k];
                            }
                        }
                    }
                    wlv = pe_best;
                    if (p.automask) {
                        const float am = p.automask[(long long)n * HW + q * W + gxr];
                        if (am <= wlv) { wlv = am; sel_q = -1; }   // mask is first in the cat: wins ties
                    }
                }
                const bool own = inside && pcol && q >= Y0 && q < Y1;
                if (own) {
                    warp_sum += wlv;
                    if (scale == p.L - 1) {
                        const long long o = (long long)n * HW + q * W + gxr;
                        if (p.viz_loss) p.viz_loss[o] = wlv;
#pragma unroll
                        for (int s = 0; s < S; ++s)
                            if (p.viz_warped[s]) {
#pragma unroll
                                for (int c = 0; c < C; ++c)
                                    p.viz_warped[s][((long long)n * C + c) * HW + q * W + gxr] = b.xm[s][c] + rc[c];
                            }
                    }
                }
                if (!BWD) {
                    // forward-only smoothness / mean-disparity sums of pixel row q (src/utils.jl:159-173)
                    const float Dr = w_dn(Db, lane);
                    float gx_ = 0.f, gy_ = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        gx_ += fabsf(b.ym[c] - w_dn(b.ym[c], lane));
                        gy_ += fabsf(b.ym[c] - cur.ym[c]);
                    }
                    if (own) {
                        if (has_right) ssx += fabsf(Db - Dr) * MD2_EXP(-gx_ * (1.0f / C));
                        if (q + 1 < H) ssy += fabsf(Db - d) * MD2_EXP(-gy_ * (1.0f / C));
                        dsum += Db;
                    }
                }
                if (BWD) {
                    const float k = (sel_q >= 0) ? up_photo * (PHOTO_ALPHA / C) * (-0.5f) : 0.f;
                    const int e0 = w_up(sel_q, lane), e2 = w_dn(sel_q, lane);
                    const float m0 = (e0 == 0) ? wl : 0.f, m1 = (sel_q == 0) ? 1.f : 0.f, m2 = (e2 == 0) ? wr : 0.f;
#pragma unroll
                    for (int j = 0; j < 3 * C; ++j) {
                        const float c1 = cf[j] * k;
                        const float c0 = w_up(c1, lane), c2 = w_dn(c1, lane);
                        rcf.t[j] = fmaf(wl, c0, fmaf(wr, c2, c1));
                        rcf.s0[j] = (S > 1) ? fmaf(m0, c0, fmaf(m2, c2, m1 * c1)) : 0.f;
                    }
                }
            }

            // ================= P(i-2): backward of pixel row i-2 =================
            if (BWD && i >= Y0 + 1) {
                const int r = i - 2;
                // vertical smoothness edge of row r (towards r+1), also needed as the "up" edge of row r+1
                float ey = 0.f;
                {
                    float g = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) g += fabsf(a.ym[c] - b.ym[c]);
                    if (r >= 0 && r + 1 < H) ey = cyn * sgnf(Da - Db) * MD2_EXP(-g * (1.0f / C));
                }
                if (i >= Y0 + 2) {
                    const bool valid = pcol;
                    const float wu = (r == 1) ? 2.f : 1.f, wd = (r == H - 2) ? 2.f : 1.f;
                    const float pyr = (float)(r + 1);
                    const int rs = slot == 2 ? 0 : slot + 1;     // ring slot of row i-2
                    float sl[NSL4 * 4];
#pragma unroll
                    for (int k = 0; k < NSL4; ++k) {
                        const Vec4 s4 = ring[(rs * RING4 + k) * 32 + lane];
                        sl[4 * k] = s4.x; sl[4 * k + 1] = s4.y; sl[4 * k + 2] = s4.z; sl[4 * k + 3] = s4.w;
                    }
                    float dbar_z = 0.f, zr = 0.f;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const Vec4 g4 = ring[(rs * RING4 + NSL4 + s) * 32 + lane];
                        zr = g4.w;
                        // d loss / d warped_s at this pixel
                        float ibar[C];
                        bool act = false;
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            float sa, sb_, sg;
                            const float ta = fmaf(wu, ra.t[3 * c], fmaf(wd, rcf.t[3 * c], rb.t[3 * c]));
                            const float tb = fmaf(wu, ra.t[3 * c + 1], fmaf(wd, rcf.t[3 * c + 1], rb.t[3 * c + 1]));
                            const float tgm = fmaf(wu, ra.t[3 * c + 2], fmaf(wd, rcf.t[3 * c + 2], rb.t[3 * c + 2]));
                            if (S == 1) { sa = ta; sb_ = tb; sg = tgm; }
                            else {
                                const float za = fmaf(wu, ra.s0[3 * c], fmaf(wd, rcf.s0[3 * c], rb.s0[3 * c]));
                                const float zb = fmaf(wu, ra.s0[3 * c + 1], fmaf(wd, rcf.s0[3 * c + 1], rb.s0[3 * c + 1]));
                                const float zg = fmaf(wu, ra.s0[3 * c + 2], fmaf(wd, rcf.s0[3 * c + 2], rb.s0[3 * c + 2]));
                                if (s == 0) { sa = za; sb_ = zb; sg = zg; }
                                else { sa = ta - za; sb_ = tb - zb; sg = tgm - zg; }
                            }
                            const float xj = a.xm[s][c], yj = a.ym[c];
                            float g = fmaf(xj, sb_, fmaf(yj, sg, sa));
                            if (sel_prev == s) g += up_photo * ((1.0f - PHOTO_ALPHA) / C) * sgnf(xj - yj);
                            ibar[c] = valid ? g : 0.f;
                            act = act || (ibar[c] != 0.f);
                        }
                        Taps tp;
                        tp.x0 = 0; tp.y0 = 0; tp.x1 = 0; tp.y1 = 0; tp.fx = 0.f; tp.fy = 0.f; tp.mx = 0.f; tp.my = 0.f;
                        if (act) {   // sources not selected anywhere in the 3x3 neighbourhood skip all of this
                            const float u = g4.x, vv = g4.y, q = g4.z;
                            tp = border_taps(u, vv, W, H);
                            float du = 0.f, dv = 0.f;
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                du = fmaf(ibar[c], sl[(s * C + c) * 2 + 0], du);
                                dv = fmaf(ibar[c], sl[(s * C + c) * 2 + 1], dv);
                            }
                            du *= tp.mx; dv *= tp.my;
                            const float cb0 = du * q, cb1 = dv * q, cb2 = -(du * u + dv * vv) * q;
                            const float ap0 = fmaf(MD2_POSE(p, pb[s] + 1), pyr, apx[s][0]);
                            const float ap1 = fmaf(MD2_POSE(p, pb[s] + 4), pyr, apx[s][1]);
                            const float ap2 = fmaf(MD2_POSE(p, pb[s] + 7), pyr, apx[s][2]);
                            dbar_z += cb0 * ap0 + cb1 * ap1 + cb2 * ap2;
                            const float t0 = cb0 * zr, t1 = cb1 * zr, t2 = cb2 * zr;
                            P0[s][0] += t0; P0[s][1] += t1; P0[s][2] += t2;
                            P1[s][0] = fmaf(t0, pyr, P1[s][0]); P1[s][1] = fmaf(t1, pyr, P1[s][1]); P1[s][2] = fmaf(t2, pyr, P1[s][2]);
                            Ph[s][0] += cb0; Ph[s][1] += cb1; Ph[s][2] += cb2;
                        }
                        // source-image gradient: scatter with vertical carry + merge with the right-hand lane
                        if (p.gsrc[s]) {
                            float* gb = p.gsrc[s] + (long long)n * p.src_ns[s];
                            const bool sval = valid && act;
                            const float w00 = (1.f - tp.fx) * (1.f - tp.fy), w01 = tp.fx * (1.f - tp.fy);
                            const float w10 = (1.f - tp.fx) * tp.fy, w11 = tp.fx * tp.fy;
                            float t0[C], t1[C];
#pragma unroll
                            for (int c = 0; c < C; ++c) { t0[c] = w00 * ibar[c]; t1[c] = w01 * ibar[c]; }
                            const int key = sval ? ((tp.y0 << 16) | tp.x0) : -2;
                            const bool have = ckey[s] >= 0;
                            const bool aligned = have && sval && key == ckey[s] + (1 << 16);
                            if (aligned) {
#pragma unroll
                                for (int c = 0; c < C; ++c) { t0[c] += car0[s][c]; t1[c] += car1[s][c]; }
                            } else if (have) {
                                float* o = gb + (((ckey[s] >> 16) + 1) * W + (ckey[s] & 0xffff));
#pragma unroll
                                for (int c = 0; c < C; ++c) {
                                    MD2_ATOMIC_ADD(o + c * HW, car0[s][c]);
                                    MD2_ATOMIC_ADD(o + c * HW + 1, car1[s][c]);
                                }
                            }
                            // my right tap is the right lane's left tap
                            const int key_r = w_dn(key, lane), key_l = w_up(key, lane);
                            const bool absorbed = sval && lane < 31 && key_r == key + 1;
                            const bool absorb = sval && lane > 0 && key_l >= 0 && key_l + 1 == key;
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                const float fl = w_up(t1[c], lane);
                                if (absorb) t0[c] += fl;
                            }
                            if (sval) {
                                float* o = gb + (tp.y0 * W + tp.x0);
#pragma unroll
                                for (int c = 0; c < C; ++c) {
                                    MD2_ATOMIC_ADD(o + c * HW, t0[c]);
                                    if (!absorbed) MD2_ATOMIC_ADD(o + c * HW + 1, t1[c]);
                                }
                                ckey[s] = key;
#pragma unroll
                                for (int c = 0; c < C; ++c) { car0[s][c] = w10 * ibar[c]; car1[s][c] = w11 * ibar[c]; }
                            } else {
                                ckey[s] = -1;
                            }
                        }
                    }
                    // depth -> disparity:  dz/dd = -a z^2
                    float gd = -p.depth_a * zr * zr * dbar_z;
                    // smoothness gradient (src/utils.jl:159-173 with the mean-normalisation of
                    // src/training.jl:64-65 folded in):  A ghat_j - B
                    {
                        const float Dr = w_dn(Da, lane);
                        float g = 0.f;
#pragma unroll
                        for (int c = 0; c < C; ++c) g += fabsf(a.ym[c] - w_dn(a.ym[c], lane));
                        const float ex = has_right ? cxn * sgnf(Da - Dr) * MD2_EXP(-g * (1.0f / C)) : 0.f;
                        const float exl = w_up(ex, lane);
                        const float gh = (ex - exl) + (ey - ey_prev);
                        gd += sA * gh - sB;
                    }
                    if (!valid) gd = 0.f;
                    if (native) {
                        if (valid) p.gdisp[scale][(long long)n * HW + r * W + gxr] = gd;
                    } else {
                        int ya0, ya1; float fyu;
                        up_taps(r, usy, dh, ya0, ya1, fyu);
                        if (ya1 == ya0) fyu = 0.f;          // clamped last row: all weight on ya0
                        if (dcur < 0) dcur = ya0;
                        while (dcur < ya0) {                // rows only advance: flush the finished one
                            flush_low(p, scale, n, dcur, da0, fxa, b0, pm, p0, pp, lane, dsm);
                            da0 = da1; da1 = 0.f; ++dcur;
                        }
                        da0 = fmaf(1.f - fyu, gd, da0);
                        da1 = fmaf(fyu, gd, da1);
                    }
                }
                ey_prev = ey;
            }
            // roll the row state
            a = b; b = cur;
            ra = rb; rb = rcf;
            Da = Db; Db = d;
            sel_prev = sel_q;
            slot = slot == 2 ? 0 : slot + 1;
        }

        if (BWD) {
            // flush the carried lower tap pairs of the last row
#pragma unroll
            for (int s = 0; s < S; ++s)
                if (p.gsrc[s] && ckey[s] >= 0) {
                    float* o = p.gsrc[s] + (long long)n * p.src_ns[s] + (((ckey[s] >> 16) + 1) * W + (ckey[s] & 0xffff));
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        MD2_ATOMIC_ADD(o + c * HW, car0[s][c]);
                        MD2_ATOMIC_ADD(o + c * HW + 1, car1[s][c]);
                    }
                }
            if (!native && dcur >= 0) {
                flush_low(p, scale, n, dcur, da0, fxa, b0, pm, p0, pp, lane, dsm);
                flush_low(p, scale, n, dcur + 1, da1, fxa, b0, pm, p0, pp, lane, dsm);
            }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;
        v[0] = warp_sum; v[1] = ssx; v[2] = ssy; v[3] = dsum;
        if (BWD) {
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
                for (int k = 0; k < 3; ++k) {   // G = sum cbar (z p)^T, p = (px, py, 1); h = sum cbar
                    v[NSTAT + 12 * s + 3 * k + 0] = px * P0[s][k];
                    v[NSTAT + 12 * s + 3 * k + 1] = P1[s][k];
                    v[NSTAT + 12 * s + 3 * k + 2] = P0[s][k];
                    v[NSTAT + 12 * s + 9 + k] = Ph[s][k];
                }
        }
    }

    // one finished low-res row of the upsample adjoint: horizontal pass across the lanes, then one
    // atomic per touched low-res element
    static MD2_DEV void flush_low(const FusedParams& p, int scale, int n, int row, float aval, float fxa, int b0,
                                  int pm, int p0, int pp, int lane, float* dsm) {
        const int dw = p.dw[scale], dh = p.dh[scale];
        dsm[lane] = (1.f - fxa) * aval;
        dsm[32 + lane] = fxa * aval;
        w_sync();
        float s = 0.f;
        for (int j = p0; j < pp; ++j) s += dsm[j];
        for (int j = pm; j < p0; ++j) s += dsm[32 + j];
        w_sync();
        const int bx = b0 + lane;
        if (s != 0.f && bx >= 0 && bx < dw && row < dh)
            MD2_ATOMIC_ADD(p.gdisp[scale] + (long long)n * dw * dh + row * dw + bx, s);
    }
};

}  // namespace md2
