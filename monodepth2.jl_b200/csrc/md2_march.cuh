// Fused view-synthesis loss, "marching warp" pipeline (the hot path).
//
// Same maths as md2_fused.cuh (src/training.jl:42-70 and its Zygote pullback), different mapping:
// ONE WARP owns a 32-column strip of one (scale, image) and marches down a chunk of rows; lane =
// image column.  Per row the warp runs three pipelined stages, all rolling state in registers:
//   L(i)    disparity -> depth -> backproject/pose/project -> 4-tap border
//           gather of the S source frames; horizontal 3-sums for the SSIM windows come from the
//           neighbouring lanes by warp shuffle
//   W(i-1)  vertical rolling 3-sums -> SSIM + L1 photometric error, arg-min over sources, automask,
//           loss partial sums; (backward) the per-window SSIM gradient coefficients and their
//           horizontal adjoint 3-sums (shuffles)
//   P(i-2)  (backward) vertical adjoint sums -> d loss / d warped, sampler / projection / depth
//           adjoints, pose accumulators, source-image scatter (lower tap pair carried to the next
//           row, right tap merged into the right-hand lane), smoothness gradient, disparity
//           gradient
// The disparity of every scale arrives at full resolution (low-res decoder scales are upsampled by
// the prep kernel into an L2-resident scratch, and their gradient is brought back by a gather-form
// adjoint kernel afterwards: md2_fused.cu), so all work items run the same code.
// Nothing is shared between warps: no block barriers, the halo is 2 columns each side (28 of 32
// lanes produce outputs) and 2 rows at the chunk ends.  The only shared memory is a per-lane
// 3-row ring holding what P(i-2) needs from L(i-2) (sampler taps, slopes, projection factors).
// The row loop is unrolled by 3 with rotating roles, so rolling the 3-row state costs no moves.
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
inline bool w_any(bool p) { return emu_ballot(p ? 1 : 0) != 0u; }
inline void w_sync() { emu_ballot(0); }
inline int w_popc(unsigned int m) { return __builtin_popcount(m); }
inline float f_sat(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
inline float f_rcp(float x) { return 1.0f / x; }
inline float f_ex2(float x) { return exp2f(x); }
inline int f_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float i_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
template <class T> inline void keep(T&) {}
inline float g_ld(const float* q) { return *q; }
inline float g_ld1(const float* q) { return q[1]; }
inline void g_st(float* q, float v) { *q = v; }
inline void g_red(float* q, float v) { *q += v; }
inline void g_red1(float* q, float v) { q[1] += v; }
inline void g_pf(const float*) {}
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
MD2_DEV bool w_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
MD2_DEV void w_sync() { __syncwarp(); }
MD2_DEV int w_popc(unsigned int m) { return __popc(m); }
MD2_DEV float f_sat(float x) { return __saturatef(x); }
MD2_DEV float f_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MD2_DEV float f_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MD2_DEV int f_as_int(float f) { return __float_as_int(f); }
MD2_DEV float i_as_float(int i) { return __int_as_float(i); }
// make a loop-invariant value opaque so that the compiler keeps it in a register instead of
// re-deriving it from kernel parameters / special registers in every row
MD2_DEV void keep(float& v) { asm volatile("" : "+f"(v)); }
MD2_DEV void keep(int& v) { asm volatile("" : "+r"(v)); }
template <class T> MD2_DEV void keep(T*& v) { asm volatile("" : "+l"(v)); }
// global-memory accesses with an explicit state space (pointers pinned by keep() have lost their
// provenance, so plain dereferences would become generic LD / ATOM with address-space checks)
MD2_DEV float g_ld(const float* q) { float v; asm("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(q)); return v; }
MD2_DEV float g_ld1(const float* q) { float v; asm("ld.global.nc.f32 %0, [%1+4];" : "=f"(v) : "l"(q)); return v; }
MD2_DEV void g_st(float* q, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(q), "f"(v)); }
MD2_DEV void g_red(float* q, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q), "f"(v)); }
MD2_DEV void g_red1(float* q, float v) { asm volatile("red.global.add.f32 [%0+4], %1;" ::"l"(q), "f"(v)); }
// pull a line into L1 one row ahead of its use (every row of the march touches new lines)
#ifndef MD2_PREFETCH
#define MD2_PREFETCH 1
#endif
MD2_DEV void g_pf(const float* q) {
#if MD2_PREFETCH == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
#elif MD2_PREFETCH == 2
    float dummy;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(dummy) : "l"(q));
#endif
}
#endif

// c * sign(d), 0 at d == 0  (abs'(0) = 0 as in ChainRules)
MD2_DEV float sgn_scaled(float d, float c) {
    return d != 0.0f ? i_as_float(f_as_int(c) ^ (f_as_int(d) & (int)0x80000000)) : 0.0f;
}

template <int C, int S, bool BWD>
struct March {
    static constexpr int HALO = BWD ? 2 : 1;
    static constexpr int OW = 32 - 2 * HALO;             // output columns per strip
    static constexpr int NPART = NSTAT + 12 * S;
    // what P(i-2) needs from L(i-2), per source: A = mx q, B = my q, u, v, fx, fy, gather offset,
    // C slopes d/dix, C slopes d/diy; plus the depth z
    static constexpr int SRCF = 7 + 2 * C;
    static constexpr int NSTF = S * SRCF + 1;
    static constexpr int NST4 = (NSTF + 3) / 4;
    static constexpr int SMEM_FLOATS = BWD ? (3 * NST4 * 32 * 4) : 4;   // the ring
    static_assert(NPART <= 32, "one lane per partial sum");

    struct Row {           // one pixel row: horizontal 3-sums (window column centred on this lane) + own values
        float hx[S][C], hxx[S][C], hxy[S][C], hy[C], hyy[C];
        float xm[S][C], ym[C];   // this lane's own (centred) warped / target values
        float D;                 // this lane's disparity
    };
    struct Coef {          // one window row: horizontal adjoint 3-sums of the coefficient maps
        float t[3 * C];    // all windows
        float s0[3 * C];   // windows whose selected source is 0
        int sel;           // this lane's selected source (-1: automask)
    };

    struct Ctx {           // loop invariants of one work item
        int W, H, HW, Y0, Y1, n, scale, lane;
        int gxm, gxr;
        bool col_img, pcol, has_right, do_viz;
        const float* tg;             // target image + this lane's column
        const float* sb[S];
        float* gb[S];
        const float* dp;             // full-resolution disparity of this (scale, image)
        float* gd;                   // its gradient
        const float* am;             // automask of this image or null
        float rc[C];
        float apx[S][3];
        int pb[S];
        float px, Wf, Hf;
        float kq;                    // wcol ? up_photo * alpha/C * (-1/2) : 0
        float cl1;                   // up_photo * (1-alpha)/C
        float mp;                    // pcol ? 1 : 0
        float wl, wr;                // horizontal reflect-pad adjoint weights of this pixel column
        float cxn, cyn, sA, sB, nega;
        Vec4* ring;
    };

    struct Acc {           // per-lane accumulators of one work item
        float warp_sum, ssx, ssy, dsum;
        float P0[S][3], P1[S][3], Ph[S][3];
        float car0[S][C], car1[S][C];
        int coff[S];
        float ey_prev;
    };

    // ---- L(i) ----
    template <int SLOT>
    static MD2_DEV void stage_load(const FusedParams& p, const Ctx& c, Acc& acc, Row& cur, int i) {
        const int lane = c.lane;
        int gym = i == -1 ? 1 : (i == c.H ? c.H - 2 : i);
        gym = gym < 0 ? 0 : (gym > c.H - 1 ? c.H - 1 : gym);
        const float py = (float)(gym + 1);
        const int toff = gym * c.W;
        const float d = g_ld(c.dp + (toff + c.gxm));
        if (gym + 1 < c.H) {   // next row's disparity / target lines
            g_pf(c.dp + (toff + c.W + c.gxm));
#pragma unroll
            for (int ch = 0; ch < C; ++ch) g_pf(c.tg + (ch * c.HW + toff + c.W));
        }
        cur.D = d;
        const float zv = rcp_acc(fmaf(d, p.depth_a, p.depth_b));
        float Tc[C], Xc[S][C];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) Tc[ch] = g_ld(c.tg + (ch * c.HW + toff)) - c.rc[ch];
        float st[NST4 * 4];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float ap0 = fmaf(MD2_POSE(p, c.pb[s] + 1), py, c.apx[s][0]);
            const float ap1 = fmaf(MD2_POSE(p, c.pb[s] + 4), py, c.apx[s][1]);
            const float ap2 = fmaf(MD2_POSE(p, c.pb[s] + 7), py, c.apx[s][2]);
            const float c0 = fmaf(zv, ap0, MD2_POSE(p, c.pb[s] + 9));
            const float c1 = fmaf(zv, ap1, MD2_POSE(p, c.pb[s] + 10));
            const float c2 = fmaf(zv, ap2, MD2_POSE(p, c.pb[s] + 11));
            const float q = rcp_acc(c2 + PROJ_EPS);
            const float u = c0 * q, vv = c1 * q;
            // border taps (NNlib grid_sample :border, align-corners); the 2x2 cell is kept inside the image
            const float cu = fminf(fmaxf(u, 1.0f), c.Wf) - 1.0f;
            const float cv = fminf(fmaxf(vv, 1.0f), c.Hf) - 1.0f;
            int x0 = (int)cu, y0 = (int)cv;
            x0 = x0 < c.W - 2 ? x0 : c.W - 2;
            y0 = y0 < c.H - 2 ? y0 : c.H - 2;
            const float fx = cu - (float)x0, fy = cv - (float)y0;
            const int off = y0 * c.W + x0;
            const float* r0 = c.sb[s] + off;
            const float* r1 = r0 + c.W;
            if (y0 + 2 < c.H) {   // the source row the next image row will need
#pragma unroll
                for (int ch = 0; ch < C; ++ch) g_pf(r1 + (ch * c.HW + c.W));
            }
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                const float v00 = g_ld(r0 + ch * c.HW), v01 = g_ld1(r0 + ch * c.HW), v10 = g_ld(r1 + ch * c.HW), v11 = g_ld1(r1 + ch * c.HW);
                const float dtop = v01 - v00, dbot = v11 - v10, dl = v10 - v00;
                const float dd = dbot - dtop;
                const float dix = fmaf(fy, dd, dtop);          // d value / d ix
                const float diy = fmaf(fx, dd, dl);            // d value / d iy
                Xc[s][ch] = fmaf(fy, diy, fmaf(fx, dtop, v00)) - c.rc[ch];
                if (BWD) { st[s * SRCF + 7 + ch] = dix; st[s * SRCF + 7 + C + ch] = diy; }
            }
            if (BWD) {
                // clip-gradient masks (0 where the un-clipped coordinate is <= 1 or >= size), folded into q
                st[s * SRCF + 0] = (u > 1.0f && u < c.Wf) ? q : 0.0f;
                st[s * SRCF + 1] = (vv > 1.0f && vv < c.Hf) ? q : 0.0f;
                st[s * SRCF + 2] = u; st[s * SRCF + 3] = vv;
                st[s * SRCF + 4] = fx; st[s * SRCF + 5] = fy;
                st[s * SRCF + 6] = i_as_float(off);
            }
        }
        if (BWD) {
            st[S * SRCF] = zv;
#pragma unroll
            for (int k = NSTF; k < NST4 * 4; ++k) st[k] = 0.f;
#pragma unroll
            for (int k = 0; k < NST4; ++k) {
                Vec4 s4; s4.x = st[4 * k]; s4.y = st[4 * k + 1]; s4.z = st[4 * k + 2]; s4.w = st[4 * k + 3];
                c.ring[(SLOT * NST4 + k) * 32 + lane] = s4;
            }
        }
        // horizontal 3-sums (window column centred on this lane)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            const float yl = w_up(Tc[ch], lane), yr = w_dn(Tc[ch], lane);
            cur.ym[ch] = Tc[ch];
            cur.hy[ch] = yl + Tc[ch] + yr;
            cur.hyy[ch] = fmaf(yr, yr, fmaf(Tc[ch], Tc[ch], yl * yl));
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float xl = w_up(Xc[s][ch], lane), xr = w_dn(Xc[s][ch], lane);
                cur.xm[s][ch] = Xc[s][ch];
                cur.hx[s][ch] = xl + Xc[s][ch] + xr;
                cur.hxx[s][ch] = fmaf(xr, xr, fmaf(Xc[s][ch], Xc[s][ch], xl * xl));
                cur.hxy[s][ch] = fmaf(xr, yr, fmaf(Xc[s][ch], Tc[ch], xl * yl));
            }
        }
    }

    // ---- W(i-1): windows centred on row q = i-1; rows a = i-2, b = i-1, cur = i ----
    static MD2_DEV void stage_windows(const FusedParams& p, const Ctx& c, Acc& acc, const Row& a, const Row& b,
                                      const Row& cur, Coef& out, int i) {
        const int lane = c.lane;
        const int q = i - 1;
        const bool row_in = q >= 0 && q < c.H;
        const bool row_own = q >= c.Y0 && q < c.Y1;
        // SSIM from 9-sample sums centred on rc, everything scaled by 81 (mu9 = 9 mu, ...):
        //   S = A B / (Cc D), A = 2 mux muy + c1, B = 2 sxy + c2, Cc = mux^2 + muy^2 + c1, D = sx + sy + c2
        constexpr float C1 = 81.0f * SSIM_C1, C2 = 81.0f * SSIM_C2;
        float pe_best = 0.f;
        int sel = 0;
        float cf[3 * C];
#pragma unroll
        for (int k = 0; k < 3 * C; ++k) cf[k] = 0.f;
        float my9[C], Y2[C], VY[C], sy[C];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            sy[ch] = a.hy[ch] + b.hy[ch] + cur.hy[ch];
            const float syy = a.hyy[ch] + b.hyy[ch] + cur.hyy[ch];
            my9[ch] = fmaf(9.0f, c.rc[ch], sy[ch]);
            Y2[ch] = fmaf(my9[ch], my9[ch], C1);
            VY[ch] = fmaf(-sy[ch], sy[ch], fmaf(9.0f, syy, C2));
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            float ssum = 0.f, lsum = 0.f;
            float cs[3 * C];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                const float sx = a.hx[s][ch] + b.hx[s][ch] + cur.hx[s][ch];
                const float sxx = a.hxx[s][ch] + b.hxx[s][ch] + cur.hxx[s][ch];
                const float sxy = a.hxy[s][ch] + b.hxy[s][ch] + cur.hxy[s][ch];
                const float mx9 = fmaf(9.0f, c.rc[ch], sx);
                const float A = fmaf(2.0f * mx9, my9[ch], C1);
                const float Cc = fmaf(mx9, mx9, Y2[ch]);
                const float B = fmaf(-2.0f * sx, sy[ch], fmaf(18.0f, sxy, C2));
                const float D = fmaf(-sx, sx, fmaf(9.0f, sxx, VY[ch]));
                const float rC = f_rcp(Cc), rD = f_rcp(D);
                const float inv = rC * rD;
                const float Sv = A * B * inv;
                const float raw = fmaf(-0.5f, Sv, 0.5f);
                const float sc = f_sat(raw);
                ssum += sc;
                lsum += fabsf(b.ym[ch] - b.xm[s][ch]);
                if (BWD) {
                    // dS/dx_j = alpha + beta x'_j + gamma y'_j for CENTRED member values x' = x - rc
                    const float pass = (sc == raw) ? 1.0f : 0.0f;     // clamp passes the gradient on [0,1]
                    const float S18 = 18.0f * Sv;
                    const float beta = -S18 * rD;
                    const float gamma = 18.0f * (A * inv);
                    const float alpha = 2.0f * fmaf(my9[ch] * (B - A), inv, Sv * mx9 * (rD - rC));
                    cs[3 * ch + 0] = fmaf(c.rc[ch], beta + gamma, alpha) * pass;
                    cs[3 * ch + 1] = beta * pass;
                    cs[3 * ch + 2] = gamma * pass;
                }
            }
            const float pe = PHOTO_ALPHA * (ssum * (1.0f / C)) + (1.0f - PHOTO_ALPHA) * (lsum * (1.0f / C));
            if (s == 0 || pe < pe_best) {   // strict <: first index wins ties (findmin)
                pe_best = pe; sel = s;
                if (BWD) {
#pragma unroll
                    for (int k = 0; k < 3 * C; ++k) cf[k] = cs[k];
                }
            }
        }
        float wlv = pe_best;
        if (c.am) {
            const int qc = q < 0 ? 0 : (q > c.H - 1 ? c.H - 1 : q);
            const float am = g_ld(c.am + (qc * c.W + c.gxm));
            if (am <= wlv) { wlv = am; sel = -1; }   // mask is first in the cat: wins ties
        }
        const bool own = row_own && c.pcol;
        acc.warp_sum += own ? wlv : 0.f;
        if (c.do_viz && own) {
            const long long o = (long long)c.n * c.HW + q * c.W + c.gxm;
            if (p.viz_loss) p.viz_loss[o] = wlv;
#pragma unroll
            for (int s = 0; s < S; ++s)
                if (p.viz_warped[s]) {
#pragma unroll
                    for (int ch = 0; ch < C; ++ch)
                        p.viz_warped[s][((long long)c.n * C + ch) * c.HW + q * c.W + c.gxm] = b.xm[s][ch] + c.rc[ch];
                }
        }
        if (!BWD) {
            // forward-only smoothness / mean-disparity sums of pixel row q (src/utils.jl:159-173)
            const float Dr = w_dn(b.D, lane);
            float gx_ = 0.f, gy_ = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                gx_ += fabsf(b.ym[ch] - w_dn(b.ym[ch], lane));
                gy_ += fabsf(b.ym[ch] - cur.ym[ch]);
            }
            if (own) {
                if (c.has_right) acc.ssx += fabsf(b.D - Dr) * MD2_EXP(-gx_ * (1.0f / C));
                if (q + 1 < c.H) acc.ssy += fabsf(b.D - cur.D) * MD2_EXP(-gy_ * (1.0f / C));
                acc.dsum += b.D;
            }
        }
        if (BWD) {
            // scale by the upstream cotangent of this window (0 outside the image / where the automask won)
            const float k = (row_in && sel >= 0) ? c.kq : 0.f;
            out.sel = sel;
            const int e0 = w_up(sel, lane), e2 = w_dn(sel, lane);
            const float m0 = (e0 == 0) ? c.wl : 0.f, m1 = (sel == 0) ? 1.f : 0.f, m2 = (e2 == 0) ? c.wr : 0.f;
#pragma unroll
            for (int j = 0; j < 3 * C; ++j) {
                const float c1 = cf[j] * k;
                const float c0 = w_up(c1, lane), c2 = w_dn(c1, lane);
                out.t[j] = fmaf(c.wl, c0, fmaf(c.wr, c2, c1));
                out.s0[j] = (S > 1) ? fmaf(m0, c0, fmaf(m2, c2, m1 * c1)) : 0.f;
            }
        }
    }

    // ---- P(i-2): backward of pixel row r = i-2; rows a = i-2, b = i-1; window rows ra = r-1, rb = r, rc = r+1 ----
    template <int RSLOT>
    static MD2_DEV void stage_pixels(const FusedParams& p, const Ctx& c, Acc& acc, const Row& a, const Row& b,
                                     const Coef& ra, const Coef& rb, const Coef& rcf, int i) {
        const int lane = c.lane;
        const int r = i - 2;
        // vertical smoothness edge of row r (towards r+1); it is also the "up" edge of row r+1
        float ey = 0.f;
        {
            float g = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) g += fabsf(a.ym[ch] - b.ym[ch]);
            const float w = f_ex2(g * (-1.4426950408889634f / C)) * c.cyn;
            if (r >= 0 && r + 1 < c.H) ey = sgn_scaled(a.D - b.D, w);
        }
        if (i >= c.Y0 + 2) {
            const float wu = (r == 1) ? 2.f : 1.f, wd = (r == c.H - 2) ? 2.f : 1.f;
            const float pyr = (float)(r + 1);
            float st[NST4 * 4];
#pragma unroll
            for (int k = 0; k < NST4; ++k) {
                const Vec4 s4 = c.ring[(RSLOT * NST4 + k) * 32 + lane];
                st[4 * k] = s4.x; st[4 * k + 1] = s4.y; st[4 * k + 2] = s4.z; st[4 * k + 3] = s4.w;
            }
            const float zr = st[S * SRCF];
            float dbar_z = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                // d loss / d warped_s at this pixel
                float ibar[C];
                bool act = false;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    float sa, sb_, sg;
                    const float ta = fmaf(wu, ra.t[3 * ch], fmaf(wd, rcf.t[3 * ch], rb.t[3 * ch]));
                    const float tb = fmaf(wu, ra.t[3 * ch + 1], fmaf(wd, rcf.t[3 * ch + 1], rb.t[3 * ch + 1]));
                    const float tgm = fmaf(wu, ra.t[3 * ch + 2], fmaf(wd, rcf.t[3 * ch + 2], rb.t[3 * ch + 2]));
                    if (S == 1) { sa = ta; sb_ = tb; sg = tgm; }
                    else {
                        const float za = fmaf(wu, ra.s0[3 * ch], fmaf(wd, rcf.s0[3 * ch], rb.s0[3 * ch]));
                        const float zb = fmaf(wu, ra.s0[3 * ch + 1], fmaf(wd, rcf.s0[3 * ch + 1], rb.s0[3 * ch + 1]));
                        const float zg = fmaf(wu, ra.s0[3 * ch + 2], fmaf(wd, rcf.s0[3 * ch + 2], rb.s0[3 * ch + 2]));
                        if (s == 0) { sa = za; sb_ = zb; sg = zg; }
                        else { sa = ta - za; sb_ = tb - zb; sg = tgm - zg; }
                    }
                    const float xj = a.xm[s][ch], yj = a.ym[ch];
                    float g = fmaf(xj, sb_, fmaf(yj, sg, sa));
                    if (rb.sel == s) g += sgn_scaled(xj - yj, c.cl1);
                    ibar[ch] = g * c.mp;
                    act = act || (ibar[ch] != 0.f);
                }
                // sources not selected anywhere in the 3x3 neighbourhood of any lane skip all of this
                if (w_any(act)) {
                    const float qa = st[s * SRCF + 0], qb = st[s * SRCF + 1], u = st[s * SRCF + 2], vv = st[s * SRCF + 3];
                    const float fx = st[s * SRCF + 4], fy = st[s * SRCF + 5];
                    const int off = f_as_int(st[s * SRCF + 6]);
                    float du = 0.f, dv = 0.f;
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) {
                        du = fmaf(ibar[ch], st[s * SRCF + 7 + ch], du);
                        dv = fmaf(ibar[ch], st[s * SRCF + 7 + C + ch], dv);
                    }
                    const float cb0 = du * qa, cb1 = dv * qb;
                    const float cb2 = -fmaf(cb0, u, cb1 * vv);
                    const float ap0 = fmaf(MD2_POSE(p, c.pb[s] + 1), pyr, c.apx[s][0]);
                    const float ap1 = fmaf(MD2_POSE(p, c.pb[s] + 4), pyr, c.apx[s][1]);
                    const float ap2 = fmaf(MD2_POSE(p, c.pb[s] + 7), pyr, c.apx[s][2]);
                    dbar_z = fmaf(cb0, ap0, fmaf(cb1, ap1, fmaf(cb2, ap2, dbar_z)));
                    const float t0 = cb0 * zr, t1 = cb1 * zr, t2 = cb2 * zr;
                    acc.P0[s][0] += t0; acc.P0[s][1] += t1; acc.P0[s][2] += t2;
                    acc.P1[s][0] = fmaf(t0, pyr, acc.P1[s][0]); acc.P1[s][1] = fmaf(t1, pyr, acc.P1[s][1]);
                    acc.P1[s][2] = fmaf(t2, pyr, acc.P1[s][2]);
                    acc.Ph[s][0] += cb0; acc.Ph[s][1] += cb1; acc.Ph[s][2] += cb2;
                    // source-image gradient: scatter with vertical carry + merge with the right-hand lane
                    if (c.gb[s]) {
                        const bool sval = act;     // (ibar is already 0 outside the output columns)
                        const float gx1 = 1.f - fx, gy1 = 1.f - fy;
                        float tq0[C], tq1[C];
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) { tq0[ch] = gx1 * gy1 * ibar[ch]; tq1[ch] = fx * gy1 * ibar[ch]; }
                        const int key = sval ? off : -2;
                        const bool have = acc.coff[s] >= 0;
                        const bool aligned = have && key == acc.coff[s] + c.W;
                        if (aligned) {
#pragma unroll
                            for (int ch = 0; ch < C; ++ch) { tq0[ch] += acc.car0[s][ch]; tq1[ch] += acc.car1[s][ch]; }
                        } else if (have) {
                            float* o = c.gb[s] + (acc.coff[s] + c.W);
#pragma unroll
                            for (int ch = 0; ch < C; ++ch) {
                                g_red(o + ch * c.HW, acc.car0[s][ch]);
                                g_red1(o + ch * c.HW, acc.car1[s][ch]);
                            }
                        }
                        // my right tap is the right lane's left tap
                        const int key_r = w_dn(key, lane), key_l = w_up(key, lane);
                        const bool absorbed = sval && lane < 31 && key_r == key + 1;
                        const bool absorb = sval && lane > 0 && key_l >= 0 && key_l + 1 == key;
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) {
                            const float fl = w_up(tq1[ch], lane);
                            if (absorb) tq0[ch] += fl;
                        }
                        if (sval) {
                            float* o = c.gb[s] + off;
#pragma unroll
                            for (int ch = 0; ch < C; ++ch) {
                                g_red(o + ch * c.HW, tq0[ch]);
                                if (!absorbed) g_red1(o + ch * c.HW, tq1[ch]);
                                acc.car0[s][ch] = gx1 * fy * ibar[ch];
                                acc.car1[s][ch] = fx * fy * ibar[ch];
                            }
                        }
                        acc.coff[s] = sval ? off : -1;
                    }
                } else if (c.gb[s]) {
                    // nobody scatters into source s on this row: flush what the previous row carried
                    if (acc.coff[s] >= 0) {
                        float* o = c.gb[s] + (acc.coff[s] + c.W);
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) {
                            g_red(o + ch * c.HW, acc.car0[s][ch]);
                            g_red1(o + ch * c.HW, acc.car1[s][ch]);
                        }
                    }
                    acc.coff[s] = -1;
                }
            }
            // depth -> disparity:  dz/dd = -a z^2
            float gd = c.nega * zr * zr * dbar_z;
            // smoothness gradient (src/utils.jl:159-173 with the mean-normalisation of
            // src/training.jl:64-65 folded in):  A ghat_j - B
            {
                const float Dr = w_dn(a.D, lane);
                float g = 0.f;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) g += fabsf(a.ym[ch] - w_dn(a.ym[ch], lane));
                const float w = f_ex2(g * (-1.4426950408889634f / C)) * c.cxn;
                const float ex = c.has_right ? sgn_scaled(a.D - Dr, w) : 0.f;
                const float exl = w_up(ex, lane);
                const float gh = (ex - exl) + (ey - acc.ey_prev);
                gd += fmaf(c.sA, gh, -c.sB);
            }
            gd *= c.mp;
            if (c.pcol) g_st(c.gd + (r * c.W + c.gxm), gd);   // (gxm == gxr on the output columns)
        }
        acc.ey_prev = ey;
    }

    template <int PH>
    static MD2_DEV void step(const FusedParams& p, const Ctx& c, Acc& acc, Row& a, Row& b, Row& cur, Coef& ra, Coef& rb,
                             Coef& rcf, int i) {
        stage_load<PH>(p, c, acc, cur, i);
        stage_windows(p, c, acc, a, b, cur, rcf, i);
        if (BWD) stage_pixels<(PH + 1) % 3>(p, c, acc, a, b, ra, rb, rcf, i);
    }

    // one warp, one (strip sx, chunk cy, scale*N+n = z) work item; on return lane-local partial
    // sums are in v[0..NPART)
    static MD2_DEV void run(const FusedParams& p, int sx, int cy, int z, int lane, float* wsm, float (&v)[32]) {
        Ctx c;
        c.lane = lane;
        c.W = p.W; c.H = p.H; c.HW = p.W * p.H;
        c.scale = z / p.N; c.n = z - c.scale * p.N;
        const int X0 = sx * OW;
        c.Y0 = cy * p.m_R;
        c.Y1 = (c.Y0 + p.m_R < c.H) ? c.Y0 + p.m_R : c.H;
        // this lane's column (reflect-pad(1): only -1 and W are ever used by an in-image window)
        c.gxr = X0 - HALO + lane;
        int gxm = c.gxr == -1 ? 1 : (c.gxr == c.W ? c.W - 2 : c.gxr);
        c.gxm = gxm < 0 ? 0 : (gxm > c.W - 1 ? c.W - 1 : gxm);
        c.col_img = c.gxr >= 0 && c.gxr < c.W;
        const bool wcol = c.col_img && lane >= 1 && lane <= 30;             // window column
        c.pcol = c.col_img && lane >= HALO && lane < 32 - HALO;             // output pixel column
        c.has_right = c.col_img && c.gxr + 1 < c.W;
        c.px = (float)(c.gxm + 1);
        c.Wf = (float)c.W; c.Hf = (float)c.H;
        c.do_viz = c.scale == p.L - 1 && (p.viz_loss != nullptr || p.viz_warped[0] != nullptr || (S > 1 && p.viz_warped[S - 1] != nullptr));

        const float* tgn = p.tgt + (long long)c.n * p.tgt_ns;
        c.tg = tgn + c.gxm;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            c.sb[s] = p.src[s] + (long long)c.n * p.src_ns[s];
            c.gb[s] = (BWD && p.gsrc[s]) ? p.gsrc[s] + (long long)c.n * p.src_ns[s] : nullptr;
        }
        c.dp = p.dfull[c.scale] + (long long)c.n * c.HW;
        c.gd = BWD ? p.gfull[c.scale] + (long long)c.n * c.HW : nullptr;
        c.am = p.automask ? p.automask + (long long)c.n * c.HW : nullptr;

        // centring constant of the window sums (any constant is exact; a local value keeps the
        // centred squares small): the target at the middle of the strip chunk
        {
            const int ym = (c.Y0 + c.Y1) >> 1;
            const int xm = X0 + OW / 2 < c.W ? X0 + OW / 2 : c.W - 1;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) c.rc[ch] = tgn[ch * c.HW + ym * c.W + xm];
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            c.pb[s] = p.pose_slot + (s * p.N + c.n) * 12;
#pragma unroll
            for (int k = 0; k < 3; ++k)   // A[:,0] px + A[:,2]: the lane-constant part of A p
                c.apx[s][k] = fmaf(MD2_POSE(p, c.pb[s] + 3 * k), c.px, MD2_POSE(p, c.pb[s] + 3 * k + 2));
        }
        const float up_photo = p.gloss * p.loss_scale / ((float)c.W * (float)c.H * (float)p.N);
        c.kq = wcol ? up_photo * (PHOTO_ALPHA / C) * (-0.5f) : 0.f;
        c.cl1 = up_photo * ((1.0f - PHOTO_ALPHA) / C);
        c.mp = c.pcol ? 1.f : 0.f;
        c.wl = (c.gxr == 1) ? 2.f : 1.f;
        c.wr = (c.gxr == c.W - 2) ? 2.f : 1.f;
        c.cxn = 1.0f / ((float)(c.W - 1) * (float)c.H * (float)p.N);
        c.cyn = 1.0f / ((float)c.W * (float)(c.H - 1) * (float)p.N);
        c.nega = -p.depth_a;
        c.sA = 0.f; c.sB = 0.f;
        if (BWD) {
            const float* st = p.stats + ((long long)c.scale * p.N + c.n) * NSTAT;
            const float up_s = p.gloss * p.loss_scale * p.smooth_w[c.scale];
            c.sA = up_s;
            if (p.normalize_disp) {
                const float m = st[3] / (float)c.HW + 1e-7f;
                c.sA = up_s / m;
                c.sB = up_s * (c.cxn * st[1] + c.cyn * st[2]) / (m * m * (float)c.HW);
            }
        }
        c.ring = reinterpret_cast<Vec4*>(wsm);
        // pin the per-lane invariants in registers (otherwise they are re-derived in every row)
        keep(c.gxm); keep(c.tg);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            keep(c.sb[s]);
            if (BWD) keep(c.gb[s]);
#pragma unroll
            for (int k = 0; k < 3; ++k) keep(c.apx[s][k]);
        }
        keep(c.dp);
        if (BWD) { keep(c.gd); keep(c.kq); keep(c.cl1); keep(c.mp); keep(c.sA); keep(c.sB); }
#pragma unroll
        for (int ch = 0; ch < C; ++ch) keep(c.rc[ch]);

        Acc acc;
        acc.warp_sum = acc.ssx = acc.ssy = acc.dsum = 0.f;
        acc.ey_prev = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            acc.coff[s] = -1;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc.P0[s][k] = acc.P1[s][k] = acc.Ph[s][k] = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) acc.car0[s][ch] = acc.car1[s][ch] = 0.f;
        }
        Row r0, r1, r2;
        Coef k0, k1, k2;
#pragma unroll
        for (int j = 0; j < 3 * C; ++j) k0.t[j] = k0.s0[j] = k1.t[j] = k1.s0[j] = k2.t[j] = k2.s0[j] = 0.f;
        k0.sel = k1.sel = k2.sel = -1;

        // rows i0, i0+1: load only; then every row runs L(i), W(i-1) [, P(i-2)]; the loop is unrolled
        // by 3 so that the three row / coefficient registers rotate roles without moves
        const int i0 = c.Y0 - HALO, iend = c.Y1 + HALO;
        stage_load<0>(p, c, acc, r0, i0);
        stage_load<1>(p, c, acc, r1, i0 + 1);
        int i = i0 + 2;
        for (; i + 2 < iend; i += 3) {
            step<2>(p, c, acc, r0, r1, r2, k0, k1, k2, i);
            step<0>(p, c, acc, r1, r2, r0, k1, k2, k0, i + 1);
            step<1>(p, c, acc, r2, r0, r1, k2, k0, k1, i + 2);
        }
        if (i < iend) {
            step<2>(p, c, acc, r0, r1, r2, k0, k1, k2, i);
            if (i + 1 < iend) step<0>(p, c, acc, r1, r2, r0, k1, k2, k0, i + 1);
        }

        if (BWD) {
            // flush the carried lower tap pairs of the last row
#pragma unroll
            for (int s = 0; s < S; ++s)
                if (c.gb[s] && acc.coff[s] >= 0) {
                    float* o = c.gb[s] + (acc.coff[s] + c.W);
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) {
                        g_red(o + ch * c.HW, acc.car0[s][ch]);
                        g_red1(o + ch * c.HW, acc.car1[s][ch]);
                    }
                }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;
        v[0] = acc.warp_sum; v[1] = acc.ssx; v[2] = acc.ssy; v[3] = acc.dsum;
        if (BWD) {
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
                for (int k = 0; k < 3; ++k) {   // G = sum cbar (z p)^T, p = (px, py, 1); h = sum cbar
                    v[NSTAT + 12 * s + 3 * k + 0] = (float)(c.gxm + 1) * acc.P0[s][k];
                    v[NSTAT + 12 * s + 3 * k + 1] = acc.P1[s][k];
                    v[NSTAT + 12 * s + 3 * k + 2] = acc.P0[s][k];
                    v[NSTAT + 12 * s + 9 + k] = acc.Ph[s][k];
                }
        }
    }
};

}  // namespace md2
