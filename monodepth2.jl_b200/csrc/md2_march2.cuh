// Fused view-synthesis loss, value + gradient: the single-warp marching kernel ("march v2", the hot path of
// md2_view_synthesis_loss_{bwd,fwdbwd}).
//
// Same maths as src/training.jl:42-70 and its Zygote pullback.  A work item is a 32-column strip of
// one (scale, image), a chunk of rows tall; lane = image column (28 output columns + 2 halo columns each side).  ONE warp
// marches down the rows and runs, in every iteration i, five software-pipelined stages that work on different rows, so
// that the instruction stream of the warp always holds several independent dependency chains (latency is hidden by
// instruction-level parallelism inside the warp; there is no inter-warp synchronisation at all):
//
//     C(i)    the 4-tap border gathers of row i (issued one iteration earlier) have returned: bilinear value and slopes
//             of the S warped sources, horizontal 3-sums for the SSIM windows (neighbouring lanes by warp shuffle)
//     A(i+1)  disparity / target of row i+1 (loaded one iteration earlier) -> depth -> backproject / pose / project ->
//             issue the gathers of row i+1 and the disparity / target loads of row i+2
//     W(i-1)  vertical 3-sums -> SSIM + L1 photometric error of the windows centred on row i-1, arg-min over sources,
//             automask, loss partial sum, SSIM gradient coefficients (alpha, beta, gamma) of the selected source,
//             smoothness gradient of the row
//     H(i-1)  horizontal adjoint 3-sums of that window row (adjoint of reflect-pad o mean-pool), routed per source
//     P(i-2)  vertical adjoint sums (rolling accumulators) -> d loss / d warped of pixel row i-2, sampler / projection /
//             depth adjoints -> disparity gradient, 12 pose accumulators per source, source-image scatter
//
// Everything that exists once per source is computed for the S = 2 sources of the reference (source_ids = [1, 3]) at
// once with packed fp32x2 instructions (FFMA2 / FADD2 / FMUL2 of sm_100: two fp32 lanes per issue slot, scalar
// operands broadcast for free), see SV<S> below.  Rows in flight between the stages live in a private shared-memory
// ring of the warp (each lane only ever reads back its own column, so no barrier of any kind is needed); the row loop
// has no divergent branches: validity of rows / columns / selections is carried by multiplicative masks and predicates.
//
// Like md2_march.cuh the code is written against the tiny warp interface of that file (w_up / w_dn / w_shfl, s_st4 /
// s_ld4, g_ld / g_st / g_red), so tests/emul runs the very same source on the CPU with 32 cooperative fibers.
#pragma once
#include "md2_march.cuh"

namespace md2 {

#ifndef MD2_M2_UNROLL
#define MD2_M2_UNROLL 1   // row-loop unroll factor of the marching kernel (1: rolled)
#endif
constexpr int M2_UNROLL = MD2_M2_UNROLL;

// ---------------------------------------------------------------------------------------------------------------
// SV<S>: one float per source frame with element-wise arithmetic; S = 2 is a packed f32x2 register pair
// ---------------------------------------------------------------------------------------------------------------
template <int S> struct SV;
template <> struct SV<1> { float v[1]; };
template <> struct alignas(8) SV<2> { float v[2]; };

#if defined(MD2_WARP_EMU)
template <int S> MD2_DEV SV<S> sv_fma(SV<S> a, SV<S> b, SV<S> c) { SV<S> r; for (int s = 0; s < S; ++s) r.v[s] = fmaf(a.v[s], b.v[s], c.v[s]); return r; }
template <int S> MD2_DEV SV<S> sv_mul(SV<S> a, SV<S> b) { SV<S> r; for (int s = 0; s < S; ++s) r.v[s] = a.v[s] * b.v[s]; return r; }
template <int S> MD2_DEV SV<S> sv_add(SV<S> a, SV<S> b) { SV<S> r; for (int s = 0; s < S; ++s) r.v[s] = a.v[s] + b.v[s]; return r; }
template <int S> MD2_DEV SV<S> sv_sub(SV<S> a, SV<S> b) { SV<S> r; for (int s = 0; s < S; ++s) r.v[s] = a.v[s] - b.v[s]; return r; }
#else
MD2_DEV unsigned long long sv_pk(const SV<2>& a) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.v[0]), "f"(a.v[1])); return r; }
MD2_DEV SV<2> sv_upk(unsigned long long r) { SV<2> a; asm("mov.b64 {%0, %1}, %2;" : "=f"(a.v[0]), "=f"(a.v[1]) : "l"(r)); return a; }
MD2_DEV SV<2> sv_fma(SV<2> a, SV<2> b, SV<2> c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(sv_pk(a)), "l"(sv_pk(b)), "l"(sv_pk(c))); return sv_upk(r); }
MD2_DEV SV<2> sv_mul(SV<2> a, SV<2> b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(sv_pk(a)), "l"(sv_pk(b))); return sv_upk(r); }
MD2_DEV SV<2> sv_add(SV<2> a, SV<2> b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(sv_pk(a)), "l"(sv_pk(b))); return sv_upk(r); }
MD2_DEV SV<2> sv_sub(SV<2> a, SV<2> b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(sv_pk(a)), "l"(sv_pk(b))); return sv_upk(r); }
MD2_DEV SV<1> sv_fma(SV<1> a, SV<1> b, SV<1> c) { SV<1> r; r.v[0] = fmaf(a.v[0], b.v[0], c.v[0]); return r; }
MD2_DEV SV<1> sv_mul(SV<1> a, SV<1> b) { SV<1> r; r.v[0] = a.v[0] * b.v[0]; return r; }
MD2_DEV SV<1> sv_add(SV<1> a, SV<1> b) { SV<1> r; r.v[0] = a.v[0] + b.v[0]; return r; }
MD2_DEV SV<1> sv_sub(SV<1> a, SV<1> b) { SV<1> r; r.v[0] = a.v[0] - b.v[0]; return r; }
#endif
template <int S> MD2_DEV SV<S> sv_bc(float a) {
    SV<S> r;
#pragma unroll
    for (int s = 0; s < S; ++s) r.v[s] = a;
    return r;
}
template <int S> MD2_DEV SV<S> sv_neg(SV<S> a) {   // (folds into the operand modifiers of FFMA2 / FADD2)
    SV<S> r;
#pragma unroll
    for (int s = 0; s < S; ++s) r.v[s] = -a.v[s];
    return r;
}
template <int S> MD2_DEV SV<S> sv_up(SV<S> a, int lane) {
    SV<S> r;
#pragma unroll
    for (int s = 0; s < S; ++s) r.v[s] = w_up(a.v[s], lane);
    return r;
}
template <int S> MD2_DEV SV<S> sv_dn(SV<S> a, int lane) {
    SV<S> r;
#pragma unroll
    for (int s = 0; s < S; ++s) r.v[s] = w_dn(a.v[s], lane);
    return r;
}
template <int S> MD2_DEV float sv_hsum(SV<S> a) {
    float r = a.v[0];
#pragma unroll
    for (int s = 1; s < S; ++s) r += a.v[s];
    return r;
}
// 1 / x: approx + one Newton step (<= 1 ulp), as rcp_acc
template <int S> MD2_DEV SV<S> sv_rcp_acc(SV<S> x) {
#if defined(MD2_WARP_EMU)
    SV<S> r;
    for (int s = 0; s < S; ++s) r.v[s] = 1.0f / x.v[s];
    return r;
#else
    SV<S> r0;
#pragma unroll
    for (int s = 0; s < S; ++s) r0.v[s] = f_rcp(x.v[s]);
    const SV<S> e = sv_fma(sv_neg(x), r0, sv_bc<S>(1.0f));
    return sv_fma(r0, e, r0);
#endif
}

MD2_DEV float f_fma_sat(float a, float b, float c) { return f_sat(fmaf(a, b, c)); }
// floor to int32, saturating, NaN -> 0 (cvt.rmi.s32.f32)
#if defined(MD2_WARP_EMU)
MD2_DEV int f_floor_i(float x) {
    if (!(x == x)) return 0;
    const float f = floorf(x);
    return f >= 2147483648.0f ? 2147483647 : (f <= -2147483648.0f ? (-2147483647 - 1) : (int)f);
}
#else
MD2_DEV int f_floor_i(float x) { return __float2int_rd(x); }
#endif

// predicated global accesses of the row loop (no branch, no reconvergence point: the predicate rides on the instruction)
#if defined(MD2_WARP_EMU)
MD2_DEV void g_st_if(float* q, float v, bool pr) { if (pr) *q = v; }
#else
MD2_DEV void g_st_if(float* q, float v, bool pr) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p st.global.f32 [%0], %1; }" ::"l"(q), "f"(v), "r"((int)pr) : "memory");
}
#endif

#ifndef MD2_M2_RC
#define MD2_M2_RC 0
#endif
#ifndef MD2_M2_HCARRY
#define MD2_M2_HCARRY 0
#endif
#ifndef MD2_M2_RED2
#define MD2_M2_RED2 0   // 0: scalar reductions; 1: vector pairs for C = 3; 2: always
#endif
#ifndef MD2_M2_MERGE
#define MD2_M2_MERGE 0
#endif
#ifndef MD2_M2_MAXREG_C1
#define MD2_M2_MAXREG_C1 128
#endif
#ifndef MD2_M2_MAXREG_C3
#define MD2_M2_MAXREG_C3 255   // (129 .. 255 registers all give two warps per scheduler: take the room, no spills)
#endif

// AM: the call has an automask map (src/training.jl:60-62); DBG: test hook, also exports the discrete decisions of every
// pixel (FusedParams::dbg, md2.h: debug_choices) -- a separate instantiation, the production kernels carry none of it
// GRAD = false: the forward-only call (md2_view_synthesis_loss_fwd, and the writer of the visualisation outputs): stages C, A and
// W alone -- no pixel packets, no adjoint stages; the loss, smoothness and mean-disparity sums of every item instead
template <int C, int S, bool AM, bool DBG = false, bool GRAD = true>
struct March2 {
    using V = SV<S>;
    static constexpr int HALO = 2;
    static constexpr int OW = 32 - 2 * HALO;             // output columns per strip
    static constexpr int NPART = NSTAT + 12 * S;
    static_assert(NPART <= 32, "one lane per partial sum");
    // ---- private shared-memory ring of the warp: one slot per row in flight, floats per lane ----
    // pixel packet (read by P two iterations after the row was consumed), source index fastest so that an SV is a
    // register pair after a 128-bit load:  u[S] v[S] off[S] fx[S] fy[S] z ym[C] | xm[C][S] dxq[C][S] dyq[C][S]
    //   (u, v: projected coordinates; off: gather offset; fx, fy: bilinear fractions; z: depth; ym / xm: centred target
    //   and warped values; dxq, dyq: slopes d warped / d(ix, iy) times the clip mask times 1 / (c3 + eps))
    static constexpr int P_U = 0, P_V = S, P_OFF = 2 * S, P_FX = 3 * S, P_FY = 4 * S, P_Z = 5 * S, P_YM = 5 * S + 1;
    static constexpr int P_XM = P_YM + C, P_DX = P_XM + S * C, P_DY = P_DX + S * C;
    static constexpr int NPF = P_DY + S * C;
    static constexpr int NE4 = (P_YM + C) / 4;           // 128-bit words that are complete once the geometry of a row is known (stage A)
    static constexpr int NP4 = (NPF + 3) / 4;            // the remaining NP4 - NE4 words are written by stage C
    // window-sum history (read by W one and two iterations later): hx[C][S] hxx[C][S] hxy[C][S] | hy[C] hyy[C]
    static constexpr int H_X = 0, H_XX = S * C, H_XY = 2 * S * C, H_Y = 3 * S * C, H_YY = H_Y + C;
    static constexpr int NHF = H_YY + C;
    static constexpr int NH4 = (NHF + 3) / 4;
    // keep the window sums of the previous row in registers (one ring load less per row and 128-bit word) where the register
    // budget has room for them: C = 1
    static constexpr bool HCARRY = MD2_M2_HCARRY != 0 && C == 1;
    static constexpr int NSLOT = 4;                      // pixel packets of rows i-2 .. i+1 are in flight
    static constexpr int NHIST = 2;                      // window sums of rows i-2, i-1 (row i takes the slot of row i-2 once that has been read)
    static constexpr int HIST0 = GRAD ? NSLOT * NP4 * 32 : 0;  // Vec4 index of the history region (forward-only: no pixel packets)
    static constexpr int TOTAL4 = (GRAD ? NSLOT * NP4 : 0) + NHIST * NH4;   // Vec4 per lane
    static constexpr int SMEM_FLOATS = TOTAL4 * 32 * 4;
    static constexpr int THREADS = 32;
    // (registers are allocated per scheduler: <= 128 -> 4 warps, <= 168 -> 3, <= 255 -> 2.  C = 3 with two sources does not fit 168
    // without spilling, so it takes the whole 2-warp budget; C = 3 with one source fits 168)
    static constexpr int MAXREG = !GRAD ? (C == 1 ? 96 : 168) : (C == 1 ? MD2_M2_MAXREG_C1 : (S == 1 ? 168 : MD2_M2_MAXREG_C3));
    static_assert(P_U + S <= NE4 * 4 && P_V + S <= NE4 * 4 && P_OFF + S <= NE4 * 4, "u, v, off must fit the early words");

    // image row read for march row i (reflect-pad(1) above and below the image, clamped beyond)
    static MD2_DEV int image_row(int i, int H) {
        const int a = i < 0 ? -i : i;
        const int r = 2 * (H - 1) - a;
        int gym = a < r ? a : r;
        gym = gym < 0 ? 0 : gym;
        return gym < H - 1 ? gym : H - 1;
    }

    struct Ctx {            // per-item invariants of a lane
        int W, H, HW, Y0, Y1, lane;
        int qlo;            // first window row of this item that exists: max(0, Y0 - 1)
        const float* tg;    // target image + this lane's column
        const float* dp;    // full-resolution disparity of this (scale, image) + this lane's column
        const float* am;    // automask of this image + this lane's column
        float* gd;          // full-resolution disparity gradient of this (scale, image) + this lane's column
        const float* sb[S];
        float* gb[S];
        int has_gb;
        float rc[C];        // centring constant of the window sums
        V epx[3], e1[3], bb[3];   // lane-constant part of E p, E[:,1], b (+ eps in bb[2])   (cam = z (p + E p) + b per source, precompose_e)
        float pxf;          // this lane's 1-based pixel column
        int gx;             // ... 0-based
        float Wf, Hf, da, db;
        float kq;           // window column valid ? up_photo * alpha / C * (-1/2) * 2 : 0   (the coefficients carry a factor 1/2)
        float cl1;          // up_photo * (1 - alpha) / C
        float mp;           // output pixel column ? 1 : 0
        float cxr, cyn;     // has_right ? 1 / ((W-1) H N) : 0,  1 / (W (H-1) N)     (the means of src/utils.jl:172)
        float wl, wr;       // horizontal reflect-pad adjoint weights of this pixel column
        float sA, sB, nega; // smoothness gradient A ghat - B (appendix A.6), -depth_a
        ring_ref rr;        // this lane's Vec4 column of the warp's ring
        int* dbg;           // DBG: decisions of this (scale, image) + this lane's column
        int do_viz;         // forward-only: this item writes the visualisation outputs (last scale, src/training.jl:71-74)
        float* vz_loss;     // ... warp-loss map of this image + this lane's column
        float* vz_warp[S];  // ... warped sources
    };

    struct Carry {          // loop-carried state
        V G[C][4];          // the four taps of row i per channel (loads issued by A(i))
        V fx, fy, qa, qb;   // bilinear fractions / clip-masked 1 / (c3 + eps) of row i
        float Tc[C], d, zc; // centred target values / disparity / depth of row i
        float Tn[C], dn;    // raw target values / disparity of row i+1 (loads issued by A(i))
        float amn;          // automask value of window row i (load issued one iteration earlier)
        int gy;             // image row of the row whose raw loads are in Tn / dn
        V xmp[C];           // warped values, target values, disparity of row i-1
        float ymp[C], Dp;
        int selp;           // selection / smoothness gradient of pixel row i-2 (from W one iteration ago)
        float ghp, ey_prev;
        V B[3 * C], Cq[3 * C];   // vertical adjoint accumulators of pixel rows i-1 and i
        V P0[3], P1[3], Ph[3];   // pose accumulators
        float warp_sum;
        float hprev[HCARRY ? NH4 * 4 : 1];   // HCARRY: window sums of row i-1 (otherwise read back from the ring like those of row i-2)
        float ssx, ssy, dsum;    // forward-only: smoothness / mean-disparity sums (src/utils.jl:159-173, src/training.jl:64)
    };

    static MD2_DEV int slot_vec(int row) { return (row & (NSLOT - 1)) * (NP4 * 32); }           // Vec4 index of the first word of a row's pixel packet
    static MD2_DEV int hist_vec(int row) { return HIST0 + (row & (NHIST - 1)) * (NH4 * 32); }   // ... of a row's window sums

    // ---- A(row): geometry of a row from the disparity / target loaded one iteration ago; issues the gathers of that row
    // and the raw loads of the next one; writes the early part of the row's pixel packet into slot `it_row` ----
    static MD2_DEV void stage_issue(const Ctx& c, Carry& k, int row, int it_row) {
        const int gym = k.gy;
        const float py = (float)(gym + 1);
        const float d = k.dn;
        k.d = d;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) k.Tc[ch] = k.Tn[ch] - c.rc[ch];
        {   // raw loads of the next row
            k.gy = image_row(row + 1, c.H);
            const int toff = k.gy * c.W;
            k.dn = g_ld(c.dp + toff);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) k.Tn[ch] = g_ld(c.tg + (ch * c.HW + toff));
        }
        const float zv = rcp_acc(fmaf(d, c.da, c.db));
        k.zc = zv;
        // Projection as a displacement (precompose_e): cam = z (p + E p) + b, so
        //   u - px = (z e0 + b0 - px w) / (z + w),  w = z e2 + b2 + eps,  e = E p
        // and the gather cell / bilinear fraction come from floor(u - px) and the integer pixel position: the sampling
        // position is accurate to float32 rounding of the displacement, not of the coordinate (hundreds of pixels).
        V e[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) e[j] = sv_fma(c.e1[j], sv_bc<S>(py), c.epx[j]);
        const V w = sv_fma(sv_bc<S>(zv), e[2], c.bb[2]);                       // (PROJ_EPS is folded into bb[2])
        const V q = sv_rcp_acc(sv_add(w, sv_bc<S>(zv)));
        const V du = sv_mul(sv_fma(sv_bc<S>(-c.pxf), w, sv_fma(sv_bc<S>(zv), e[0], c.bb[0])), q);
        const V dv = sv_mul(sv_fma(sv_bc<S>(-py), w, sv_fma(sv_bc<S>(zv), e[1], c.bb[1])), q);
        const V u = sv_add(du, sv_bc<S>(c.pxf)), v = sv_add(dv, sv_bc<S>(py));   // full coordinates (adjoint of the perspective divide)
        // border taps (NNlib grid_sample :border, align-corners); the 2x2 cell is kept inside the image.  Inside the image
        // 1 < u < W, i.e. 1 - px < du < W - px; beyond, the coordinate is clipped to the border (and its gradient masked)
        const float xlo = 1.0f - c.pxf, xhi = c.Wf - c.pxf, ylo = 1.0f - py, yhi = c.Hf - py;
        const int cxl = -c.gx, cxh = c.W - 2 - c.gx, cyl = -gym, cyh = c.H - 2 - gym;   // cell range relative to this pixel
        V flx, fly, adx, ady;
        int off[S], x0s[S], y0s[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int fxi = f_floor_i(du.v[s]), fyi = f_floor_i(dv.v[s]);       // (saturating conversion; NaN -> 0)
            const int rx = fxi < cxl ? cxl : (fxi > cxh ? cxh : fxi), ry = fyi < cyl ? cyl : (fyi > cyh ? cyh : fyi);
            const int x0 = c.gx + rx, y0 = gym + ry;                            // 0-based gather cell, inside the image
            flx.v[s] = (float)fxi; fly.v[s] = (float)fyi;
            adx.v[s] = (float)(fxi - rx); ady.v[s] = (float)(fyi - ry);         // 0 inside; < 0 / > 0 where clipped left / right
            x0s[s] = x0; y0s[s] = y0;
            off[s] = y0 * c.W + x0;
            // clip-gradient masks (0 where the un-clipped coordinate is <= 1 or >= size), folded into q
            k.qa.v[s] = (du.v[s] > xlo && du.v[s] < xhi) ? q.v[s] : 0.0f;
            k.qb.v[s] = (dv.v[s] > ylo && dv.v[s] < yhi) ? q.v[s] : 0.0f;
            const float* r0 = c.sb[s] + off[s];
            const float* r1 = c.sb[s] + (off[s] + c.W);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                k.G[ch][0].v[s] = g_ld(r0 + ch * c.HW); k.G[ch][1].v[s] = g_ld1(r0 + ch * c.HW);
                k.G[ch][2].v[s] = g_ld(r1 + ch * c.HW); k.G[ch][3].v[s] = g_ld1(r1 + ch * c.HW);
            }
        }
        {   // bilinear fractions: the fraction of the displacement inside the image, 0 / 1 where the coordinate is clipped
            const V sx_ = sv_add(sv_sub(du, flx), adx), sy_ = sv_add(sv_sub(dv, fly), ady);
#pragma unroll
            for (int s = 0; s < S; ++s) { k.fx.v[s] = f_sat(sx_.v[s]); k.fy.v[s] = f_sat(sy_.v[s]); }
        }
        if (DBG) {
            if ((unsigned)(row - c.Y0) < (unsigned)(c.Y1 - c.Y0) && c.mp != 0.f) {
#pragma unroll
                for (int s = 0; s < S; ++s)
                    c.dbg[(row * c.W) * (1 + S) + 1 + s] = x0s[s] | (y0s[s] << 14) | (k.qa.v[s] != 0.0f || (q.v[s] == 0.0f && du.v[s] > xlo && du.v[s] < xhi) ? (1 << 29) : 0) |
                                                            (k.qb.v[s] != 0.0f || (q.v[s] == 0.0f && dv.v[s] > ylo && dv.v[s] < yhi) ? (1 << 30) : 0);
            }
        }
        if (GRAD && NE4 > 0) {   // early words of the pixel packet (the rest is carried in registers until stage C stores it)
            float pk[NE4 > 0 ? NE4 * 4 : 4];
#pragma unroll
            for (int j = 0; j < NE4 * 4; ++j) pk[j] = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                if (P_U + s < NE4 * 4) pk[P_U + s] = u.v[s];
                if (P_V + s < NE4 * 4) pk[P_V + s] = v.v[s];
                if (P_OFF + s < NE4 * 4) pk[P_OFF + s] = i_as_float(off[s]);
                if (P_FX + s < NE4 * 4) pk[P_FX + s] = k.fx.v[s];
                if (P_FY + s < NE4 * 4) pk[P_FY + s] = k.fy.v[s];
            }
            if (P_Z < NE4 * 4) pk[P_Z] = zv;
#pragma unroll
            for (int ch = 0; ch < C; ++ch)
                if (P_YM + ch < NE4 * 4) pk[P_YM + ch] = k.Tc[ch];
            const int base = slot_vec(it_row);
#pragma unroll
            for (int w4 = 0; w4 < NE4; ++w4) {
                Vec4 q4; q4.x = pk[4 * w4]; q4.y = pk[4 * w4 + 1]; q4.z = pk[4 * w4 + 2]; q4.w = pk[4 * w4 + 3];
                s_st4(c.rr, base + w4 * 32, q4);
            }
        }
    }

    // ---- one iteration: C(i), A(i+1), W(i-1), H(i-1), P(i-2);  it = i - (Y0 - HALO) ----
    // (the shared-memory accesses are volatile asm statements: they keep their source order, so the loads are placed
    // well ahead of their uses by hand)
    static MD2_DEV void step(const Ctx& c, Carry& k, int i, int it) {
        const int lane = c.lane;
        // window-sum history of rows i-2, i-1 (for W)
        float ha[NH4 * 4], hb[NH4 * 4];
        {
            const int sa = hist_vec(it - 2), sb = hist_vec(it - 1);
#pragma unroll
            for (int w4 = 0; w4 < NH4; ++w4) {
                const Vec4 a4 = s_ld4(c.rr, sa + w4 * 32);
                ha[4 * w4] = a4.x; ha[4 * w4 + 1] = a4.y; ha[4 * w4 + 2] = a4.z; ha[4 * w4 + 3] = a4.w;
                if (HCARRY) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) hb[4 * w4 + j] = k.hprev[4 * w4 + j];
                } else {
                    const Vec4 b4 = s_ld4(c.rr, sb + w4 * 32);
                    hb[4 * w4] = b4.x; hb[4 * w4 + 1] = b4.y; hb[4 * w4 + 2] = b4.z; hb[4 * w4 + 3] = b4.w;
                }
            }
        }
        // =========================== C(i) ===========================
        V X[C];
        float curT[C];
        const float curD = k.d;
        {
            float pk[NP4 * 4];
#pragma unroll
            for (int j = 0; j < NP4 * 4; ++j) pk[j] = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                const V v00 = k.G[ch][0], v01 = k.G[ch][1], v10 = k.G[ch][2], v11 = k.G[ch][3];
                const V dtop = sv_sub(v01, v00), dbot = sv_sub(v11, v10), dl = sv_sub(v10, v00);
                const V dd = sv_sub(dbot, dtop);
                const V dix = sv_fma(k.fy, dd, dtop);          // d value / d ix
                const V diy = sv_fma(k.fx, dd, dl);            // d value / d iy
                X[ch] = sv_sub(sv_fma(k.fy, diy, sv_fma(k.fx, dtop, v00)), sv_bc<S>(c.rc[ch]));
                const V dxq = sv_mul(dix, k.qa), dyq = sv_mul(diy, k.qb);
                curT[ch] = k.Tc[ch];
                pk[P_YM + ch] = curT[ch];
#pragma unroll
                for (int s = 0; s < S; ++s) { pk[P_XM + ch * S + s] = X[ch].v[s]; pk[P_DX + ch * S + s] = dxq.v[s]; pk[P_DY + ch * S + s] = dyq.v[s]; }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) { pk[P_FX + s] = k.fx.v[s]; pk[P_FY + s] = k.fy.v[s]; }
            pk[P_Z] = k.zc;
            const int base = slot_vec(it);
            if (GRAD) {
#pragma unroll
                for (int w4 = NE4; w4 < NP4; ++w4) {   // (the early words were written by A one iteration ago)
                    Vec4 q4; q4.x = pk[4 * w4]; q4.y = pk[4 * w4 + 1]; q4.z = pk[4 * w4 + 2]; q4.w = pk[4 * w4 + 3];
                    s_st4(c.rr, base + w4 * 32, q4);
                }
            }
        }
        // horizontal 3-sums (window column centred on this lane)
        V hx[C], hxx[C], hxy[C];
        float hy[C], hyy[C];
        {
            float hs[NH4 * 4];
#pragma unroll
            for (int j = 0; j < NH4 * 4; ++j) hs[j] = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                const float Tc = curT[ch];
                const float yl = w_up(Tc, lane), yr = w_dn(Tc, lane);
                hy[ch] = yl + Tc + yr;
                hyy[ch] = fmaf(yr, yr, fmaf(Tc, Tc, yl * yl));
                const V xl = sv_up(X[ch], lane), xr = sv_dn(X[ch], lane);
                hx[ch] = sv_add(sv_add(xl, X[ch]), xr);
                hxx[ch] = sv_fma(xr, xr, sv_fma(X[ch], X[ch], sv_mul(xl, xl)));
                hxy[ch] = sv_fma(xr, sv_bc<S>(yr), sv_fma(X[ch], sv_bc<S>(Tc), sv_mul(xl, sv_bc<S>(yl))));
                hs[H_Y + ch] = hy[ch]; hs[H_YY + ch] = hyy[ch];
#pragma unroll
                for (int s = 0; s < S; ++s) { hs[H_X + ch * S + s] = hx[ch].v[s]; hs[H_XX + ch * S + s] = hxx[ch].v[s]; hs[H_XY + ch * S + s] = hxy[ch].v[s]; }
            }
            const int sc = hist_vec(it);   // (= the slot of row i-2, which was read at the top of this iteration)
#pragma unroll
            for (int w4 = 0; w4 < NH4; ++w4) {
                Vec4 c4; c4.x = hs[4 * w4]; c4.y = hs[4 * w4 + 1]; c4.z = hs[4 * w4 + 2]; c4.w = hs[4 * w4 + 3];
                s_st4(c.rr, sc + w4 * 32, c4);
            }
            if (HCARRY) {
#pragma unroll
                for (int j = 0; j < NH4 * 4; ++j) k.hprev[j] = hs[j];
            }
        }
        // pixel packet of row i-2 (for P)
        float pk[NP4 * 4];
        if (GRAD) {
            const int base = slot_vec(it - 2);
#pragma unroll
            for (int w4 = 0; w4 < NP4; ++w4) {
                const Vec4 t = s_ld4(c.rr, base + w4 * 32);
                pk[4 * w4] = t.x; pk[4 * w4 + 1] = t.y; pk[4 * w4 + 2] = t.z; pk[4 * w4 + 3] = t.w;
            }
        }
        // =========================== A(i+1) ===========================
        const float am_q = AM ? k.amn : 0.f;
        stage_issue(c, k, i + 1, it + 1);
        const int q = i - 1;
        if (AM) {   // automask value of the next window row
            const int qn = i < 0 ? 0 : (i > c.H - 1 ? c.H - 1 : i);
            k.amn = g_ld(c.am + qn * c.W);
        }

        // =========================== W(i-1): windows centred on row q ===========================
        const bool row_in = (unsigned)(q - c.qlo) < (unsigned)(c.H - c.qlo);
        const bool row_own = (unsigned)(q - c.Y0) < (unsigned)(c.Y1 - c.Y0);
        // SSIM from 9-sample sums centred on rc, everything scaled by 81 (mu9 = 9 mu, ...):
        //   S = A B / (Cc D), A = 2 mux muy + c1, B = 2 sxy + c2, Cc = mux^2 + muy^2 + c1, D = sx + sy + c2;  Bn = -B, Dn = -D
        constexpr float C1 = 81.0f * SSIM_C1, C2 = 81.0f * SSIM_C2;
        V ssum, lsum;
        V cs[3 * C];
        V passv = sv_bc<S>(1.f);
        int dbgw = 0;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            const float sy = ha[H_Y + ch] + hb[H_Y + ch] + hy[ch];
            const float syy = ha[H_YY + ch] + hb[H_YY + ch] + hyy[ch];
            const float rc9 = 9.0f * c.rc[ch];
            const float my9 = rc9 + sy;
            const float Y2 = fmaf(my9, my9, C1);
            const float nVY = fmaf(sy, sy, fmaf(-9.0f, syy, -C2));
            V sx, sxx, sxy, bx, bxx, bxy;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                sx.v[s] = ha[H_X + ch * S + s]; sxx.v[s] = ha[H_XX + ch * S + s]; sxy.v[s] = ha[H_XY + ch * S + s];
                bx.v[s] = hb[H_X + ch * S + s]; bxx.v[s] = hb[H_XX + ch * S + s]; bxy.v[s] = hb[H_XY + ch * S + s];
            }
            sx = sv_add(sv_add(sx, bx), hx[ch]);
            sxx = sv_add(sv_add(sxx, bxx), hxx[ch]);
            sxy = sv_add(sv_add(sxy, bxy), hxy[ch]);
            const V mx9 = sv_add(sx, sv_bc<S>(rc9));
            const V A = sv_fma(mx9, sv_bc<S>(2.0f * my9), sv_bc<S>(C1));
            const V Cc = sv_fma(mx9, mx9, sv_bc<S>(Y2));
            const V Bn = sv_fma(sx, sv_bc<S>(2.0f * sy), sv_fma(sxy, sv_bc<S>(-18.0f), sv_bc<S>(-C2)));
            const V Dn = sv_fma(sx, sx, sv_fma(sxx, sv_bc<S>(-9.0f), sv_bc<S>(nVY)));
            const V den = sv_mul(Cc, Dn);
            V inv;
#pragma unroll
            for (int s = 0; s < S; ++s) inv.v[s] = f_rcp(den.v[s]);        // 1 / (Cc Dn) = -1 / (Cc D)
            const V Sv = sv_mul(sv_mul(A, Bn), inv);
            V sc, pm;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                sc.v[s] = f_fma_sat(-0.5f, Sv.v[s], 0.5f);
                pm.v[s] = (fabsf(Sv.v[s]) <= 1.0f) ? 1.f : 0.f;            // clamp passes the gradient on [0,1]
            }
            const V df = sv_sub(k.xmp[ch], sv_bc<S>(k.ymp[ch]));
            V ad;
#pragma unroll
            for (int s = 0; s < S; ++s) ad.v[s] = fabsf(df.v[s]);
            if (ch == 0) { ssum = sc; lsum = ad; }
            else { ssum = sv_add(ssum, sc); lsum = sv_add(lsum, ad); }
            if (DBG) {
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    dbgw |= (pm.v[s] != 0.f ? 1 : 0) << (2 + s * C + ch);
                    dbgw |= (df.v[s] > 0.f ? 1 : (df.v[s] < 0.f ? 2 : 0)) << (8 + 2 * (s * C + ch));
                }
            }
            if (!GRAD) continue;
            // dS/dx_j = alpha + beta x'_j + gamma y'_j for CENTRED member values x' = x - rc; all three carry a factor 1/2
            const V rDn = sv_mul(inv, Cc), rC = sv_mul(inv, Dn);
            V beta = sv_mul(sv_mul(Sv, sv_bc<S>(9.0f)), rDn);
            V gamma = sv_mul(sv_mul(A, inv), sv_bc<S>(-9.0f));
            const V T1 = sv_mul(sv_add(Bn, A), inv);
            const V T2 = sv_mul(sv_mul(Sv, mx9), sv_add(rDn, rC));
            V alpha = sv_fma(sv_bc<S>(my9), T1, sv_neg(T2));
            if (C > 1) { alpha = sv_mul(alpha, pm); beta = sv_mul(beta, pm); gamma = sv_mul(gamma, pm); }
            else passv = pm;
            cs[3 * ch + 0] = sv_fma(sv_bc<S>(c.rc[ch]), sv_add(beta, gamma), alpha);
            cs[3 * ch + 1] = beta;
            cs[3 * ch + 2] = gamma;
        }
        // photometric error per source, arg-min (strict <: first index wins ties = findmin), automask (mask wins ties)
        float pe[S];
#pragma unroll
        for (int s = 0; s < S; ++s) pe[s] = fmaf(ssum.v[s], PHOTO_ALPHA / C, lsum.v[s] * ((1.0f - PHOTO_ALPHA) / C));
        float wlv = pe[0];
        int sel = 0;
#pragma unroll
        for (int s = 1; s < S; ++s)
            if (pe[s] < wlv) { wlv = pe[s]; sel = s; }
        if (AM) {
            if (am_q <= wlv) { wlv = am_q; sel = -1; }
        }
        k.warp_sum += row_own ? wlv * c.mp : 0.f;
        if (!GRAD) {
            // ---- forward-only: visualisation outputs of the last scale, smoothness / mean-disparity sums of pixel row q ----
            const bool own = row_own && c.mp != 0.f;
            if (c.do_viz) {   // (warp-uniform: a property of the item)
                if (c.vz_loss) g_st_if(c.vz_loss + q * c.W, wlv, own);
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (c.vz_warp[s]) {
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) g_st_if(c.vz_warp[s] + (ch * c.HW + q * c.W), k.xmp[ch].v[s] + c.rc[ch], own);
                    }
            }
            const float Dr = w_dn(k.Dp, lane);
            float gxs = 0.f, gys = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                gxs += fabsf(k.ymp[ch] - w_dn(k.ymp[ch], lane));
                gys += fabsf(k.ymp[ch] - curT[ch]);
            }
            const float wx = f_ex2(gxs * (-1.4426950408889634f / C)), wy = f_ex2(gys * (-1.4426950408889634f / C));
            k.ssx += (own && c.cxr != 0.f) ? fabsf(k.Dp - Dr) * wx : 0.f;                       // (cxr != 0: the pixel has a right neighbour)
            k.ssy += (own && (unsigned)q < (unsigned)(c.H - 1)) ? fabsf(k.Dp - curD) * wy : 0.f;
            k.dsum += own ? k.Dp : 0.f;
            k.Dp = curD;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) { k.xmp[ch] = X[ch]; k.ymp[ch] = curT[ch]; }
            return;
        }
        // coefficients of the selected source, scaled by the upstream cotangent of this window
        float wp[3 * C];
        {
            const float kk = (row_in && sel >= 0) ? c.kq : 0.f;
            float f[S];
#pragma unroll
            for (int s = 0; s < S; ++s) f[s] = (sel == s) ? kk * passv.v[s] : 0.f;
#pragma unroll
            for (int j = 0; j < 3 * C; ++j) {
                float a = cs[j].v[0] * f[0];
#pragma unroll
                for (int s = 1; s < S; ++s) a = fmaf(cs[j].v[s], f[s], a);
                wp[j] = a;
            }
        }
        // smoothness gradient of pixel row q before the mean-normalisation (P applies A ghat - B):
        //   ghat = (ex(q) - ex(q)[left lane]) + (ey(q) - ey(q-1)), e = sign(d - d') exp(-mean_c |T - T'|) / count
        float gh;
        {
            const float Dr = w_dn(k.Dp, lane);
            float gxs = 0.f, gys = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                gxs += fabsf(k.ymp[ch] - w_dn(k.ymp[ch], lane));
                gys += fabsf(k.ymp[ch] - curT[ch]);
            }
            const float wx = f_ex2(gxs * (-1.4426950408889634f / C)) * c.cxr;
            const float wy = f_ex2(gys * (-1.4426950408889634f / C)) * (((unsigned)q < (unsigned)(c.H - 1)) ? c.cyn : 0.f);
            const float ex = sgn_scaled(k.Dp - Dr, wx);
            const float ey = sgn_scaled(k.Dp - curD, wy);
            const float exl = w_up(ex, lane);
            gh = (ex - exl) + (ey - k.ey_prev);
            k.ey_prev = ey;
            if (DBG) {
                const float dx_ = k.Dp - Dr, dy_ = k.Dp - curD;
                dbgw |= (sel + 1) | ((dx_ > 0.f ? 1 : (dx_ < 0.f ? 2 : 0)) << 20) | ((dy_ > 0.f ? 1 : (dy_ < 0.f ? 2 : 0)) << 22);
                if (row_own && c.mp != 0.f) c.dbg[(q * c.W) * (1 + S)] = dbgw;
            }
        }
        // =========================== H(i-1): horizontal adjoint 3-sums, routed per source ===========================
        V hw[3 * C];
        {
            const int e0 = w_up(sel, lane), e2 = w_dn(sel, lane);
            V mL, mO, mR;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                mL.v[s] = (e0 == s) ? c.wl : 0.f;
                mO.v[s] = (sel == s) ? 1.f : 0.f;
                mR.v[s] = (e2 == s) ? c.wr : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 3 * C; ++j) {
                const float c1 = wp[j];
                const float c0 = w_up(c1, lane), c2 = w_dn(c1, lane);
                hw[j] = sv_fma(mL, sv_bc<S>(c0), sv_fma(mR, sv_bc<S>(c2), sv_mul(mO, sv_bc<S>(c1))));
            }
        }
        // =========================== P(i-2): pixel row r ===========================
        const int r = i - 2;
        {
            // window row q adds to pixel rows q-1 (x2 if q is the last image row: adjoint of the reflect-pad), q, q+1 (x2 if q == 0)
            const float wup = (q == c.H - 1) ? 2.f : 1.f, wdn = (q == 0) ? 2.f : 1.f;
            V ib[3 * C];
#pragma unroll
            for (int j = 0; j < 3 * C; ++j) {
                ib[j] = sv_fma(sv_bc<S>(wup), hw[j], k.B[j]);
                k.B[j] = sv_add(k.Cq[j], hw[j]);
                k.Cq[j] = sv_mul(sv_bc<S>(wdn), hw[j]);
            }
            const bool r_own = (unsigned)(r - c.Y0) < (unsigned)(c.Y1 - c.Y0);
            const float m = r_own ? c.mp : 0.f;
            const float pyr = (float)(r + 1);
            const float zr = pk[P_Z];
            V u, v, fx, fy, cl1s;
            int off[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                u.v[s] = pk[P_U + s]; v.v[s] = pk[P_V + s]; fx.v[s] = pk[P_FX + s]; fy.v[s] = pk[P_FY + s]; off[s] = f_as_int(pk[P_OFF + s]);
                cl1s.v[s] = (k.selp == s) ? c.cl1 : 0.f;
            }
            // d loss / d warped at this pixel
            V ibar[C];
            V cb0, cb1;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                V xj, dxq, dyq;
#pragma unroll
                for (int s = 0; s < S; ++s) { xj.v[s] = pk[P_XM + ch * S + s]; dxq.v[s] = pk[P_DX + ch * S + s]; dyq.v[s] = pk[P_DY + ch * S + s]; }
                const float yj = pk[P_YM + ch];
                V gv = sv_fma(xj, ib[3 * ch + 1], sv_fma(sv_bc<S>(yj), ib[3 * ch + 2], ib[3 * ch]));
                const V df = sv_sub(xj, sv_bc<S>(yj));
                V l1;
#pragma unroll
                for (int s = 0; s < S; ++s) l1.v[s] = sgn_scaled(df.v[s], cl1s.v[s]);
                gv = sv_mul(sv_add(gv, l1), sv_bc<S>(m));
                ibar[ch] = gv;
                if (ch == 0) { cb0 = sv_mul(gv, dxq); cb1 = sv_mul(gv, dyq); }
                else { cb0 = sv_fma(gv, dxq, cb0); cb1 = sv_fma(gv, dyq, cb1); }
            }
            const V ncb2 = sv_fma(cb0, u, sv_mul(cb1, v));    // -cbar_3
            V ap[3];                                            // A p = p + E p
#pragma unroll
            for (int j = 0; j < 3; ++j) ap[j] = sv_fma(c.e1[j], sv_bc<S>(pyr), c.epx[j]);
            ap[0] = sv_add(ap[0], sv_bc<S>(c.pxf)); ap[1] = sv_add(ap[1], sv_bc<S>(pyr)); ap[2] = sv_add(ap[2], sv_bc<S>(1.0f));
            const V dz = sv_sub(sv_fma(cb0, ap[0], sv_mul(cb1, ap[1])), sv_mul(ncb2, ap[2]));
            const float dbar_z = sv_hsum(dz);
            const float zp = zr * pyr;
            k.P0[0] = sv_fma(cb0, sv_bc<S>(zr), k.P0[0]); k.P1[0] = sv_fma(cb0, sv_bc<S>(zp), k.P1[0]); k.Ph[0] = sv_add(k.Ph[0], cb0);
            k.P0[1] = sv_fma(cb1, sv_bc<S>(zr), k.P0[1]); k.P1[1] = sv_fma(cb1, sv_bc<S>(zp), k.P1[1]); k.Ph[1] = sv_add(k.Ph[1], cb1);
            k.P0[2] = sv_fma(ncb2, sv_bc<S>(-zr), k.P0[2]); k.P1[2] = sv_fma(ncb2, sv_bc<S>(-zp), k.P1[2]); k.Ph[2] = sv_sub(k.Ph[2], ncb2);
            // source-image gradient: the four taps of every pixel (predicated atomics)
            if (c.has_gb) {
                const V gx1 = sv_sub(sv_bc<S>(1.f), fx), gy1 = sv_sub(sv_bc<S>(1.f), fy);
                bool act[S];
#pragma unroll
                for (int s = 0; s < S; ++s) act[s] = false;
                V t0[C], t1[C], b0[C], b1[C];
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const V wl_ = sv_mul(gx1, ibar[ch]), wr_ = sv_mul(fx, ibar[ch]);
                    t0[ch] = sv_mul(wl_, gy1); t1[ch] = sv_mul(wr_, gy1); b0[ch] = sv_mul(wl_, fy); b1[ch] = sv_mul(wr_, fy);
#pragma unroll
                    for (int s = 0; s < S; ++s) act[s] = act[s] || (ibar[ch].v[s] != 0.f);
                }
                // Merging a pixel's right taps into the right-hand lane's left taps by shuffle (two reductions per pixel and
                // source instead of four, where neighbouring lanes sample neighbouring cells) trades 2 REDs for 2 SHFLs per
                // channel and tap row.  C = 1: measured slower (62.2 vs 59.5 us at 416x128x8: the kernel is short of issue
                // slots and LSU wavefronts, both cost the same).  C = 3: the warp has slots to spare and waits on its own
                // reductions instead -- MD2_M2_MERGE selects (0: never, 1: C = 3 only, 2: always)
                if (MD2_M2_MERGE == 2 || (MD2_M2_MERGE == 1 && C == 3)) {
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const int offa = act[s] ? off[s] : (-0x40000000 - lane);       // an idle lane neither gives nor receives
                        const int nb = w_dn(offa, lane);                                // left-tap offset of the right-hand lane
                        const bool give = act[s] && lane < 31 && nb == off[s] + 1;
                        float* o = c.gb[s] + off[s];
                        float* o1 = c.gb[s] + (off[s] + c.W);
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) {
                            float rt = w_up(give ? t1[ch].v[s] : 0.f, lane), rb = w_up(give ? b1[ch].v[s] : 0.f, lane);
                            if (lane == 0) { rt = 0.f; rb = 0.f; }                     // (shfl.up hands lane 0 its own value back)
                            if (act[s]) {
                                g_red(o + ch * c.HW, t0[ch].v[s] + rt);
                                g_red(o1 + ch * c.HW, b0[ch].v[s] + rb);
                                if (!give) { g_red1(o + ch * c.HW, t1[ch].v[s]); g_red1(o1 + ch * c.HW, b1[ch].v[s]); }
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (act[s]) {   // (ptxas turns predicated atomics into one branch each: one region per source instead)
                            float* o = c.gb[s] + off[s];
                            float* o1 = c.gb[s] + (off[s] + c.W);
#pragma unroll
                            for (int ch = 0; ch < C; ++ch) {
                                if (MD2_M2_RED2 == 2 || (MD2_M2_RED2 == 1 && C == 3)) {   // 64-bit vector reductions where the tap pair is aligned
                                    g_red_pair(o + ch * c.HW, t0[ch].v[s], t1[ch].v[s]);
                                    g_red_pair(o1 + ch * c.HW, b0[ch].v[s], b1[ch].v[s]);
                                } else {
                                    g_red(o + ch * c.HW, t0[ch].v[s]); g_red1(o + ch * c.HW, t1[ch].v[s]);
                                    g_red(o1 + ch * c.HW, b0[ch].v[s]); g_red1(o1 + ch * c.HW, b1[ch].v[s]);
                                }
                            }
                        }
                    }
                }
            }
            // depth -> disparity (dz/dd = -a z^2) + smoothness gradient with the mean-normalisation folded in (A ghat - B)
            const float gdv = fmaf(c.nega * zr * zr, dbar_z, fmaf(c.sA, k.ghp, -c.sB));
            g_st_if(c.gd + r * c.W, gdv, m != 0.f);
        }
        // carry
        k.selp = sel; k.ghp = gh; k.Dp = curD;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) { k.xmp[ch] = X[ch]; k.ymp[ch] = curT[ch]; }
    }

    // one work item; on return v[] holds this lane's partial sums (loss sum, pose sums)
    static MD2_DEV void run(const FusedParams& p, int sx, int cy, int z, int lane, float* wsm, float (&v)[32]) {
        Ctx c;
        c.lane = lane;
        c.W = p.W; c.H = p.H; c.HW = p.W * p.H;
        const int scale = z / p.N;
        const int n = z - scale * p.N;
        c.Y0 = cy * p.m_R;
        c.Y1 = (c.Y0 + p.m_R < c.H) ? c.Y0 + p.m_R : c.H;
        c.qlo = c.Y0 - 1 > 0 ? c.Y0 - 1 : 0;
        // this lane's column (reflect-pad(1): only -1 and W are ever used by an in-image window)
        const int gxr = sx * OW - HALO + lane;
        int gxm = gxr == -1 ? 1 : (gxr == c.W ? c.W - 2 : gxr);
        gxm = gxm < 0 ? 0 : (gxm > c.W - 1 ? c.W - 1 : gxm);
        const bool col_img = gxr >= 0 && gxr < c.W;
        const bool pcol = col_img && lane >= HALO && lane < 32 - HALO;      // output pixel column
        const bool wcol = col_img && lane >= 1 && lane <= 30;               // window column
        const bool has_right = col_img && gxr + 1 < c.W;
        const float* tgn = p.tgt + (long long)n * p.tgt_ns;
        c.tg = tgn + gxm;
        c.dp = p.dfull[scale] + (long long)n * c.HW + gxm;
        c.gd = p.gfull[scale] + (long long)n * c.HW + gxm;
        c.am = AM ? p.automask + (long long)n * c.HW + gxm : tgn;
        c.has_gb = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            c.sb[s] = p.src[s] + (long long)n * p.src_ns[s];
            c.gb[s] = p.gsrc[s] ? p.gsrc[s] + (long long)n * p.src_ns[s] : nullptr;
            if (p.gsrc[s]) c.has_gb = 1;
        }
        {   // Centring constant of the window sums (any constant is exact; it keeps the centred squares small).  It must not
            // depend on the work item: windows on the border between two items are computed by both, and only with the same
            // constant are the two computations bit-identical -- otherwise a near-tie of the two sources' photometric errors can
            // be decided differently by the two items, and the window's gradient is routed to source 0 in one pixel row and
            // to source 1 in the next.  MD2_M2_RC: 0 = mean of nine target samples of the image, 1 = middle of the strip chunk
            // (the item-dependent choice, for comparison only)
#if MD2_M2_RC == 1
            const int ym = (c.Y0 + c.Y1) >> 1;
            const int xm = sx * OW + OW / 2 < c.W ? sx * OW + OW / 2 : c.W - 1;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) c.rc[ch] = tgn[ch * c.HW + ym * c.W + xm];
#else
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                float acc = 0.f;
#pragma unroll
                for (int j = 1; j <= 3; ++j)
#pragma unroll
                    for (int i = 1; i <= 3; ++i) acc += tgn[ch * c.HW + (j * c.H / 4) * c.W + i * c.W / 4];
                c.rc[ch] = acc * (1.0f / 9.0f);
            }
#endif
        }
        const float px = (float)(gxm + 1);
        c.pxf = px; c.gx = gxm;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float* eb = p.pose_e + (long long)(s * p.N + n) * 12;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                c.epx[j].v[s] = fmaf(eb[3 * j], px, eb[3 * j + 2]);    // E[:,0] px + E[:,2]: the lane-constant part of E p
                c.e1[j].v[s] = eb[3 * j + 1];
                c.bb[j].v[s] = eb[9 + j] + (j == 2 ? PROJ_EPS : 0.0f);
            }
        }
        c.Wf = (float)c.W; c.Hf = (float)c.H;
        c.da = p.depth_a; c.db = p.depth_b;
        const float up_photo = p.gloss * p.loss_scale / ((float)c.W * (float)c.H * (float)p.N);
        c.kq = wcol ? up_photo * (PHOTO_ALPHA / C) * (-1.0f) : 0.f;     // (-1/2) x 2: the coefficients carry a factor 1/2
        c.cl1 = up_photo * ((1.0f - PHOTO_ALPHA) / C);
        c.mp = pcol ? 1.f : 0.f;
        const float cxn = 1.0f / ((float)(c.W - 1) * (float)c.H * (float)p.N);
        c.cxr = has_right ? cxn : 0.f;
        c.cyn = 1.0f / ((float)c.W * (float)(c.H - 1) * (float)p.N);
        c.nega = -p.depth_a;
        c.wl = (gxr == 1) ? 2.f : 1.f;
        c.wr = (gxr == c.W - 2) ? 2.f : 1.f;
        {
            const float up_s = p.gloss * p.loss_scale * p.smooth_w[scale];
            c.sA = up_s; c.sB = 0.f;
            if (GRAD && p.normalize_disp) {
                float ssx, ssy, dsum;
                prep_stats(p, scale, n, lane, ssx, ssy, dsum);
                const float m = dsum / (float)c.HW + 1e-7f;
                c.sA = up_s / m;
                c.sB = up_s * (cxn * ssx + c.cyn * ssy) / (m * m * (float)c.HW);
            }
        }
        c.rr = ring_ref_of(wsm, lane);
        c.dbg = DBG ? p.dbg + ((long long)(scale * p.N + n) * c.HW + gxm) * (1 + S) : nullptr;
        c.do_viz = 0; c.vz_loss = nullptr;
#pragma unroll
        for (int s = 0; s < S; ++s) c.vz_warp[s] = nullptr;
        if (!GRAD && scale == p.L - 1) {
            c.vz_loss = p.viz_loss ? p.viz_loss + (long long)n * c.HW + gxm : nullptr;
            c.do_viz = c.vz_loss != nullptr;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                c.vz_warp[s] = p.viz_warped[s] ? p.viz_warped[s] + (long long)n * C * c.HW + gxm : nullptr;
                if (c.vz_warp[s]) c.do_viz = 1;
            }
        }
        // pin the per-lane invariants in registers (otherwise they are re-derived from launch parameters / special
        // registers inside the row loop, with a scoreboard wait each)
        keep(c.lane); keep(c.W); keep(c.H); keep(c.rr); keep(c.mp); keep(c.kq); keep(c.da); keep(c.db);
        keep(c.tg); keep(c.dp); keep(c.gd); keep(c.wl); keep(c.wr); keep(c.sA); keep(c.sB); keep(c.cxr); keep(c.cyn);
        if (AM) keep(c.am);
#pragma unroll
        for (int s = 0; s < S; ++s) { keep(c.sb[s]); keep(c.gb[s]); }
#pragma unroll
        for (int ch = 0; ch < C; ++ch) keep(c.rc[ch]);
        {   // the ring starts out as zeros: warm-up rows read slots that this item has not written yet
            Vec4 z4; z4.x = z4.y = z4.z = z4.w = 0.f;
#pragma unroll 4
            for (int w4 = 0; w4 < TOTAL4; ++w4) s_st4(c.rr, w4 * 32, z4);
        }
        Carry k;
#pragma unroll
        for (int j = 0; j < 3 * C; ++j) { k.B[j] = sv_bc<S>(0.f); k.Cq[j] = sv_bc<S>(0.f); }
#pragma unroll
        for (int j = 0; j < 3; ++j) { k.P0[j] = sv_bc<S>(0.f); k.P1[j] = sv_bc<S>(0.f); k.Ph[j] = sv_bc<S>(0.f); }
#pragma unroll
        for (int ch = 0; ch < C; ++ch) { k.xmp[ch] = sv_bc<S>(0.f); k.ymp[ch] = 0.f; }
        k.Dp = 0.f; k.selp = -1; k.ghp = 0.f; k.ey_prev = 0.f; k.warp_sum = 0.f; k.zc = 0.f; k.amn = 0.f;
        k.ssx = 0.f; k.ssy = 0.f; k.dsum = 0.f;
#pragma unroll
        for (int j = 0; j < (HCARRY ? NH4 * 4 : 1); ++j) k.hprev[j] = 0.f;
        const int i0 = c.Y0 - (GRAD ? HALO : 1), iend = c.Y1 + (GRAD ? HALO : 1);   // (forward-only: windows reach one row out)
        {   // prime the pipeline: raw loads of row i0, then A(i0)
            k.gy = image_row(i0, c.H);
            const int toff = k.gy * c.W;
            k.dn = g_ld(c.dp + toff);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) k.Tn[ch] = g_ld(c.tg + (ch * c.HW + toff));
            stage_issue(c, k, i0, 0);
            if (AM) {
                const int q0 = i0 - 1, qn = q0 < 0 ? 0 : (q0 > c.H - 1 ? c.H - 1 : q0);
                k.amn = g_ld(c.am + qn * c.W);
            }
        }
#pragma unroll M2_UNROLL
        for (int i = i0, it = 0; i < iend; ++i, ++it) step(c, k, i, it);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
        v[0] = k.warp_sum;
        if (!GRAD) { v[1] = k.ssx; v[2] = k.ssy; v[3] = k.dsum; return; }
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int j = 0; j < 3; ++j) {   // G = sum cbar (z p)^T, p = (px, py, 1); h = sum cbar
                v[NSTAT + 12 * s + 3 * j + 0] = px * k.P0[j].v[s];
                v[NSTAT + 12 * s + 3 * j + 1] = k.P1[j].v[s];
                v[NSTAT + 12 * s + 3 * j + 2] = k.P0[j].v[s];
                v[NSTAT + 12 * s + 9 + j] = k.Ph[j].v[s];
            }
    }
};

}  // namespace md2
