// Fused view-synthesis loss, "marching warp" pipeline (the hot path).
//
// Same maths as src/training.jl:42-70 and its Zygote pullback.  A work item is a 32-column strip
// of one (scale, image), a chunk of rows tall; lane = image column, the warps march down the rows
// with all rolling state in registers.  The backward runs as a two-warp producer / consumer pair:
//
//   warp F (forward)   per row i, software-pipelined over rows so that no load is consumed in the
//   iteration that issues it:
//     C(i)    consume the 4-tap border gathers of row i (issued one iteration earlier): bilinear value +
//             slopes of the S warped sources; horizontal 3-sums for the SSIM windows come from the
//             neighbouring lanes by warp shuffle
//     A(i+1)  disparity / target of row i+1 (loaded one iteration earlier) -> depth -> backproject / pose /
//             project -> issue the gathers of row i+1 and the disparity / target loads of row i+2
//     W(i-1)  vertical rolling 3-sums -> SSIM + L1 photometric error, arg-min over sources,
//             automask, loss partial sums; the per-window SSIM gradient coefficients and their
//             horizontal adjoint 3-sums (shuffles)
//     -> fills one slot of a shared-memory ring per row: the pixel packet of row i (early part: geometry,
//        written by A(i); late part: own values and slopes, written by C(i)) and the window packet of row i-1
//   warp B (backward)  per pixel row r, once slot r+2 is full:
//     P(r)    vertical adjoint sums -> d loss / d warped, sampler / projection / depth adjoints,
//             pose accumulators, source-image scatter (lower tap pair carried to the next row,
//             right tap merged into the right-hand lane), smoothness gradient, disparity gradient
//   The two warps are coupled only by full/empty mbarriers per ring slot, so the gather
//   latency of F overlaps the arithmetic of B, and each warp carries half of the register state
//   (twice the resident warps per SM of a single-warp formulation).
//
// Forward-only calls run warp F alone.  The halo is 2 columns each side (28 of 32 lanes produce
// outputs) and 2 rows at the chunk ends; F's row loop is unrolled by 3 with rotating roles, so
// rolling the 3-row state costs no moves.  The disparity of every scale arrives at full resolution
// (low-res decoder scales are upsampled by the prep kernel into an L2-resident scratch, and their
// gradient is brought back by a gather-form adjoint kernel afterwards: md2_fused.cu).
//
// The code is written against a tiny warp/block interface (w_up / w_dn / w_shfl / w_any /
// mb_wait / mb_arrive) so that tests/emul can run the very same source on the CPU with 32
// cooperative fibers per warp (MD2_WARP_EMU, test infrastructure only).
#pragma once
#include <string.h>

#include "md2_fused.cuh"

namespace md2 {

struct alignas(16) Vec4 { float x, y, z, w; };

// source-gradient scatter of warp B.  0: both tap rows of a pixel are issued at once, the right taps merged into the
// right-hand lane's left taps (~2.2 atomics per pixel and source); 1: the lower tap pair is additionally carried in
// registers to the next row (~1.6 atomics, but ~20 more instructions per row and source).  Measured at 416x128x8:
// 71.9 us (0) vs 75.6 us (1): the kernel is bound by issue slots, not by the atomics.
#ifndef MD2_SCATTER_CARRY
#define MD2_SCATTER_CARRY 0
#endif
// which warp forms the smoothness gradient of a pixel (balances the pair): 0 = warp F (sends ghat with the window
// packet), 1 = warp B (reads the target values / disparity of row r+1 from the next slot)
#ifndef MD2_SMOOTH_IN_B
#define MD2_SMOOTH_IN_B 0   // measured at 416x128x8: 72.3 us (0) vs 73.9 us (1)
#endif
#ifndef MD2_SKIP_IDLE_SOURCE
#define MD2_SKIP_IDLE_SOURCE 1   // warp B skips the adjoint of a source that no lane selected around this row (warp vote)
#endif
#ifndef MD2_PIN_MORE_C3
#define MD2_PIN_MORE_C3 0
#endif
#ifndef MD2_SCATTER_MERGE
#define MD2_SCATTER_MERGE 1   // merge a pixel's right taps into the right-hand lane's left taps (warp shuffles)
#endif

#if defined(MD2_WARP_EMU)
#define MD2_DEV inline
// emu_xchg / emu_ballot / emu_bar: tests/emul/warp_emu.h, included before this file
inline float w_shfl(float v, int src, int) { unsigned int u; memcpy(&u, &v, 4); u = emu_xchg(u, src); float r; memcpy(&r, &u, 4); return r; }
inline int w_shfl(int v, int src, int) { return (int)emu_xchg((unsigned int)v, src); }
inline float w_up(float v, int lane) { return w_shfl(v, lane > 0 ? lane - 1 : 0, lane); }
inline float w_dn(float v, int lane) { return w_shfl(v, lane < 31 ? lane + 1 : 31, lane); }
inline int w_up(int v, int lane) { return w_shfl(v, lane > 0 ? lane - 1 : 0, lane); }
inline int w_dn(int v, int lane) { return w_shfl(v, lane < 31 ? lane + 1 : 31, lane); }
inline bool w_any(bool p) { return emu_ballot(p ? 1 : 0) != 0u; }
typedef unsigned long long mbar_t;
typedef mbar_t* bar_ref;          // the mbarrier array
typedef float* ring_ref;          // the ring of packets
inline bar_ref bar_ref_of(float* q) { return reinterpret_cast<mbar_t*>(q); }
inline ring_ref ring_ref_of(float* q, int lane) { return q + 4 * lane; }   // + this lane's Vec4 column
inline void mb_arrive(bar_ref bars, int idx) { emu_mb_arrive(bars, idx); }
inline void mb_wait(bar_ref bars, int idx, int parity) { emu_mb_wait(bars, idx, parity); }
inline void s_st4(ring_ref r, int vec_index, const Vec4& v) { reinterpret_cast<Vec4*>(r)[vec_index] = v; }
inline Vec4 s_ld4(ring_ref r, int vec_index) { return reinterpret_cast<const Vec4*>(r)[vec_index]; }
inline float f_sat(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
inline float f_rcp(float x) { return 1.0f / x; }
inline float f_ex2(float x) { return exp2f(x); }
inline int f_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float i_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
template <class T> inline void keep(T&) {}
#define MD2_RARE_BLOCK() ((void)0)
inline float g_ld(const float* q) { return *q; }
inline float g_ld1(const float* q) { return q[1]; }
inline void g_st(float* q, float v) { *q = v; }
inline void g_red(float* q, float v) { *q += v; }
inline void g_red1(float* q, float v) { q[1] += v; }
inline void g_red_pair(float* q, float a, float b) { q[0] += a; q[1] += b; }
inline void g_pf(const float*) {}
#define MD2_POSE(p, idx) ((p).pose_ab[(idx)])
#else
#define MD2_DEV __device__ __forceinline__
// pre-composed pose rows (A | b per source and image) of the forward-only marching kernel: warp-uniform loads from the
// ctx's pose table (L1-resident; one table per ctx, so calls on different ctxs / streams never share state)
#define MD2_POSE(p, idx) (__ldg((p).pose_ab + (idx)))
MD2_DEV float w_shfl(float v, int src, int) { return __shfl_sync(0xffffffffu, v, src); }
MD2_DEV int w_shfl(int v, int src, int) { return __shfl_sync(0xffffffffu, v, src); }
MD2_DEV float w_up(float v, int) { return __shfl_up_sync(0xffffffffu, v, 1); }
MD2_DEV float w_dn(float v, int) { return __shfl_down_sync(0xffffffffu, v, 1); }
MD2_DEV int w_up(int v, int) { return __shfl_up_sync(0xffffffffu, v, 1); }
MD2_DEV int w_dn(int v, int) { return __shfl_down_sync(0xffffffffu, v, 1); }
MD2_DEV bool w_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
// shared-memory mbarriers (32 arrivals per phase: every lane of the signalling warp arrives).
// Named bar.sync/bar.arrive barriers would also do, but 2 x depth of them per block cap the
// resident blocks per SM at 4 (16 hardware barriers per block are reserved).
typedef unsigned long long mbar_t;
// shared-memory objects are addressed by their 32-bit shared-window offsets (no generic-address round trips)
typedef unsigned int bar_ref;
typedef unsigned int ring_ref;
MD2_DEV unsigned int smem_u32(const void* q) { return (unsigned int)__cvta_generic_to_shared(q); }
MD2_DEV bar_ref bar_ref_of(float* q) { return smem_u32(q); }
MD2_DEV ring_ref ring_ref_of(float* q, int lane) { return smem_u32(q) + 16u * lane; }   // + this lane's Vec4 column
MD2_DEV void mb_init(bar_ref bars, int idx, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8u * idx), "r"(count) : "memory");
}
MD2_DEV void mb_arrive(bar_ref bars, int idx) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8u * idx) : "memory");
}
#ifndef MD2_B_UNROLL3
#define MD2_B_UNROLL3 0   // 1: warp B row loop unrolled by 3 (no register moves for the rotating window rows); measured 2 % slower (instruction cache)
#endif
#ifndef MD2_WAIT_HINT_NS
#define MD2_WAIT_HINT_NS 20000
#endif
MD2_DEV void mb_wait(bar_ref bars, int idx, int parity) {   // returns once the phase of that parity has completed
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"   // suspends in hardware up to the time hint
        "@p bra MB_DONE_%=;\n"
        "bra MB_WAIT_%=;\n"
        "MB_DONE_%=:\n"
        "}\n" ::"r"(bars + 8u * idx), "r"(parity), "r"(MD2_WAIT_HINT_NS) : "memory");
}
MD2_DEV void s_st4(ring_ref r, int vec_index, const Vec4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(r + 16u * vec_index), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
MD2_DEV Vec4 s_ld4(ring_ref r, int vec_index) {
    Vec4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(r + 16u * vec_index) : "memory");
    return v;
}
MD2_DEV float f_sat(float x) { return __saturatef(x); }
MD2_DEV float f_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MD2_DEV float f_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MD2_DEV int f_as_int(float f) { return __float_as_int(f); }
MD2_DEV float i_as_float(int i) { return __int_as_float(i); }
// make a loop-invariant value opaque so that the compiler keeps it in a register instead of
// re-deriving it from kernel parameters / special registers in every row
MD2_DEV void keep(float& v) { asm volatile("" : "+f"(v)); }
MD2_DEV void keep(int& v) { asm volatile("" : "+r"(v)); }
MD2_DEV void keep(unsigned int& v) { asm volatile("" : "+r"(v)); }
template <class T> MD2_DEV void keep(T*& v) { asm volatile("" : "+l"(v)); }
// first statement of a warp-uniform, rarely taken block: keeps the compiler from if-converting it into predicated
// instructions that are issued on every row (a uniform branch costs two)
#define MD2_RARE_BLOCK() asm volatile("")
// global-memory accesses with an explicit state space (pointers pinned by keep() have lost their
// provenance, so plain dereferences would become generic LD / ATOM with address-space checks)
MD2_DEV float g_ld(const float* q) { float v; asm("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(q)); return v; }
MD2_DEV float g_ld1(const float* q) { float v; asm("ld.global.nc.f32 %0, [%1+4];" : "=f"(v) : "l"(q)); return v; }
MD2_DEV void g_st(float* q, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(q), "f"(v)); }
#if defined(MD2_FAKE_RED)   // timing experiment only (wrong results): plain stores in place of the reductions
MD2_DEV void g_red(float* q, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(q), "f"(v)); }
MD2_DEV void g_red1(float* q, float v) { asm volatile("st.global.f32 [%0+4], %1;" ::"l"(q), "f"(v)); }
#else
MD2_DEV void g_red(float* q, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q), "f"(v)); }
MD2_DEV void g_red1(float* q, float v) { asm volatile("red.global.add.f32 [%0+4], %1;" ::"l"(q), "f"(v)); }
#endif
// q[0] += a, q[1] += b: one 64-bit vector reduction where the pair is 8-byte aligned (half the sector operations of two
// scalar ones), two scalar reductions otherwise -- three predicated instructions, no branch
MD2_DEV void g_red_pair(float* q, float a, float b) {
    asm volatile("{ .reg .pred p; .reg .b64 t; and.b64 t, %0, 4; setp.eq.b64 p, t, 0;\n"
                 "  @p red.global.add.v2.f32 [%0], {%1, %2};\n"
                 "  @!p red.global.add.f32 [%0], %1;\n"
                 "  @!p red.global.add.f32 [%0+4], %2; }" ::"l"(q), "f"(a), "f"(b));
}
// pull a line into L1 one row ahead of its use (every row of the march touches new lines)
#ifndef MD2_PREFETCH
#define MD2_PREFETCH 0   // measured: with the register pipeline of warp F the L1 prefetch of source rows costs more than it saves
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

// smoothness / mean-disparity sums (Sx, Sy, sum d) of one (scale, image), needed by the backward
// (src/training.jl:64-65 couples all pixels of an image): the saved statistics of an earlier forward,
// or -- fused fwd+bwd call -- the prep kernel's per-block partials added in a fixed order
// (lane-strided, then a butterfly: deterministic).  Called by all 32 lanes.
MD2_DEV void prep_stats(const FusedParams& p, int scale, int n, int lane, float& ssx, float& ssy, float& dsum) {
    if (!p.prep_part) {
        const float* st = p.stats + ((long long)scale * p.N + n) * NSTAT;
        ssx = st[1]; ssy = st[2]; dsum = st[3];
        return;
    }
    const Vec4* pp = reinterpret_cast<const Vec4*>(p.prep_part) + ((long long)scale * p.N + n) * p.prep_nblk;
    float a = 0.f, b = 0.f, c = 0.f;
    for (int k = lane; k < p.prep_nblk; k += 32) {
        const Vec4 q = pp[k];
        a += q.x; b += q.y; c += q.z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += w_shfl(a, lane ^ o, lane); b += w_shfl(b, lane ^ o, lane); c += w_shfl(c, lane ^ o, lane);
    }
    ssx = a; ssy = b; dsum = c;
}

constexpr int MARCH_DEPTH = 5;       // ring slots between warp F and warp B (B holds 3, F fills the 4th and pre-fills the 5th)
constexpr int BAR_FULL = 0;          // mbarrier indices: full[s] = BAR_FULL + s, empty[s] = BAR_EMPTY + s
constexpr int BAR_EMPTY = MARCH_DEPTH;
constexpr int MARCH_NBAR = 2 * MARCH_DEPTH;
constexpr int SLOT_WRAP = 2 * MARCH_DEPTH;   // a running slot counter = slot index + MARCH_DEPTH * (fill parity)

// position in the ring: slot index and the parity of the number of times the ring has wrapped
struct Slot { int idx, par; };
MD2_DEV Slot slot_of(int counter) { Slot t; t.par = counter >= MARCH_DEPTH ? 1 : 0; t.idx = counter - t.par * MARCH_DEPTH; return t; }
MD2_DEV int counter_of(Slot t) { return t.idx + t.par * MARCH_DEPTH; }
MD2_DEV Slot next_slot(Slot t) {
    Slot n;
    const bool wrap = t.idx + 1 == MARCH_DEPTH;
    n.idx = wrap ? 0 : t.idx + 1;
    n.par = wrap ? t.par ^ 1 : t.par;
    return n;
}

template <int C, int S, bool BWD>
struct March {
    static constexpr int HALO = BWD ? 2 : 1;
    // pin more loop invariants in registers (saves re-reading %tid / launch parameters in the row loops); only where the
    // register budget has room: with C = 3 it turns into spills
    static constexpr bool PIN_MORE = (C == 1) || (MD2_PIN_MORE_C3 != 0);
    static constexpr int OW = 32 - 2 * HALO;             // output columns per strip
    static constexpr int NPART = NSTAT + 12 * S;
    // ---- ring slot layout (floats per lane) ----
    // pixel packet of row i, early part (known once the geometry of the row is): own target values ym[C],
    // disparity D, depth z, then per source: q = 1/(c3 + eps), u, v, fx, fy, gather offset
    static constexpr int O_YM = 0, O_D = C, O_Z = C + 1, O_SRC = C + 2;
    static constexpr int SRCF = 6;
    static constexpr int NEF = O_SRC + S * SRCF;
    static constexpr int NE4 = (NEF + 3) / 4;
    // late part (known once the gathers have returned): own warped values xm[S][C], then per source
    // C slopes d/dix, C slopes d/diy
    static constexpr int O_XM = NE4 * 4, O_SL = O_XM + S * C;
    static constexpr int NLF = 3 * S * C;
    static constexpr int NL4 = (NLF + 3) / 4;
    static constexpr int NPP4 = NE4 + NL4;
    static constexpr int NYD4 = (C + 1 + 3) / 4;         // Vec4s holding ym[C], D
    // window packet of row i-1: the SSIM gradient coefficients (alpha, beta, gamma)[C] of the selected source,
    // scaled by the window's upstream cotangent, the selected source, and the un-normalised smoothness
    // gradient ghat of the pixel (divergence of the edge-weighted sign field, src/utils.jl:159-173)
    static constexpr int NWPF = 3 * C + 2;
    static constexpr int NWP4 = (NWPF + 3) / 4;
    static constexpr int SLOT4 = NPP4 + NWP4;            // Vec4 per lane and slot
    static constexpr int RING_FLOATS = BWD ? (MARCH_DEPTH * SLOT4 * 32 * 4) : 0;
    static constexpr int SMEM_FLOATS = RING_FLOATS + 2 * MARCH_NBAR + 4;   // ring, then the mbarriers (8 B each)
    static constexpr int THREADS = BWD ? 64 : 32;
    static_assert(NPART <= 32, "one lane per partial sum");

    struct Row {           // one pixel row: horizontal 3-sums (window column centred on this lane) + own values
        float hx[S][C], hxx[S][C], hxy[S][C], hy[C], hyy[C];
        float xm[S][C], ym[C];   // this lane's own (centred) warped / target values
        float D;                 // this lane's disparity
    };

    struct Geo {           // what both warps need to know about the work item
        int W, H, HW, Y0, Y1, n, scale, lane, gxm, gxr;
        bool col_img, pcol, has_right;
    };

    static MD2_DEV void geometry(const FusedParams& p, int sx, int cy, int z, int lane, Geo& g) {
        g.lane = lane;
        g.W = p.W; g.H = p.H; g.HW = p.W * p.H;
        g.scale = z / p.N; g.n = z - g.scale * p.N;
        g.Y0 = cy * p.m_R;
        g.Y1 = (g.Y0 + p.m_R < g.H) ? g.Y0 + p.m_R : g.H;
        // this lane's column (reflect-pad(1): only -1 and W are ever used by an in-image window)
        g.gxr = sx * OW - HALO + lane;
        int gxm = g.gxr == -1 ? 1 : (g.gxr == g.W ? g.W - 2 : g.gxr);
        g.gxm = gxm < 0 ? 0 : (gxm > g.W - 1 ? g.W - 1 : gxm);
        g.col_img = g.gxr >= 0 && g.gxr < g.W;
        g.pcol = g.col_img && lane >= HALO && lane < 32 - HALO;             // output pixel column
        g.has_right = g.col_img && g.gxr + 1 < g.W;
    }

    // =====================================================================================
    // warp F
    // =====================================================================================
    struct CtxF {
        Geo g;
        bool do_viz;
        const float* tg;             // target image + this lane's column
        const float* sb[S];
        const float* dp;             // full-resolution disparity of this (scale, image) + this lane's column
        const float* am;             // automask of this image or null
        int has_am;                  // (kept in a register: the null test of the pointer would re-load the launch parameter every row)
        float rc[C];
        float apx[S][3];
        int pb[S];
        float Wf, Hf;
        float kq;                    // wcol ? up_photo * alpha/C * (-1/2) : 0
        float cxn, cyn;              // 1 / ((W-1) H N), 1 / (W (H-1) N): the means of src/utils.jl:172
        ring_ref ring;               // + this lane's Vec4 column
        bar_ref bars;
        Slot fill;                   // slot of the row being consumed (C / W stages); A pre-fills the next one
    };
    struct AccF { float warp_sum, ssx, ssy, dsum, ey_prev; };
    // rows in flight between the stages of warp F
    struct PipeF {
        float G[S][C][4];            // the four taps of row i per source and channel (loads issued by A(i))
        float fx[S], fy[S];          // their bilinear fractions
        float Tc[C], d;              // centred target values / disparity of row i
        float Tn[C], dn;             // raw target values / disparity of row i+1 (loads issued by A(i))
        int gy;                      // image row of the row whose raw loads are in Tn / dn
    };

    // image row read for march row i (reflect-pad(1) above and below the image, clamped beyond)
    static MD2_DEV int image_row(int i, int H) {
        // reflect about row 0 and row H-1 (-1 -> 1, H -> H-2; rows further out are never used by a window of the image),
        // then clamp (tiny images)
        const int a = i < 0 ? -i : i;
        const int r = 2 * (H - 1) - a;
        int gym = a < r ? a : r;
        gym = gym < 0 ? 0 : gym;
        return gym < H - 1 ? gym : H - 1;
    }

    static MD2_DEV int slot_vec(Slot t) { return t.idx * (SLOT4 * 32); }   // Vec4 index of a slot's first packet

    // ---- A(i): geometry of row i from the disparity / target loaded one row ago; issues the gathers of
    // row i and the disparity / target loads of row i+1; BWD: pre-fills the early part of slot `t` ----
    static MD2_DEV void stage_issue(const FusedParams& p, const CtxF& c, PipeF& f, int i, Slot t) {
        const Geo& g = c.g;
        const int gym = f.gy;
        const float py = (float)(gym + 1);
        const float d = f.dn;
        f.d = d;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) f.Tc[ch] = f.Tn[ch] - c.rc[ch];
        {   // next row
            f.gy = image_row(i + 1, g.H);
            const int toff = f.gy * g.W;
            f.dn = g_ld(c.dp + toff);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) f.Tn[ch] = g_ld(c.tg + (ch * g.HW + toff));
        }
        const float zv = rcp_acc(fmaf(d, p.depth_a, p.depth_b));
        float ek[NE4 * 4];
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
            x0 = x0 < g.W - 2 ? x0 : g.W - 2;
            y0 = y0 < g.H - 2 ? y0 : g.H - 2;
            const float fx = cu - (float)x0, fy = cv - (float)y0;
            const int off = y0 * g.W + x0;
            const float* r0 = c.sb[s] + off;
            const float* r1 = r0 + g.W;
            if (y0 + 2 < g.H) {   // the source row the next image row will need
#pragma unroll
                for (int ch = 0; ch < C; ++ch) g_pf(r1 + (ch * g.HW + g.W));
            }
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                f.G[s][ch][0] = g_ld(r0 + ch * g.HW); f.G[s][ch][1] = g_ld1(r0 + ch * g.HW);
                f.G[s][ch][2] = g_ld(r1 + ch * g.HW); f.G[s][ch][3] = g_ld1(r1 + ch * g.HW);
            }
            f.fx[s] = fx; f.fy[s] = fy;
            if (BWD) {
                ek[O_SRC + s * SRCF + 0] = q;
                ek[O_SRC + s * SRCF + 1] = u; ek[O_SRC + s * SRCF + 2] = vv;
                ek[O_SRC + s * SRCF + 3] = fx; ek[O_SRC + s * SRCF + 4] = fy;
                ek[O_SRC + s * SRCF + 5] = i_as_float(off);
            }
        }
        if (BWD) {
            ek[O_Z] = zv; ek[O_D] = d;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) ek[O_YM + ch] = f.Tc[ch];
#pragma unroll
            for (int k = NEF; k < NE4 * 4; ++k) ek[k] = 0.f;
            mb_wait(c.bars, BAR_EMPTY + t.idx, t.par ^ 1);   // the previous fill of this slot has been consumed
            const int base = slot_vec(t);
#pragma unroll
            for (int k = 0; k < NE4; ++k) {
                Vec4 s4; s4.x = ek[4 * k]; s4.y = ek[4 * k + 1]; s4.z = ek[4 * k + 2]; s4.w = ek[4 * k + 3];
                s_st4(c.ring, base + k * 32, s4);
            }
        }
    }

    // ---- C(i): the gathers of row i have returned: bilinear value + slopes (BWD: late part of the slot),
    // horizontal 3-sums (window column centred on this lane) ----
    static MD2_DEV void stage_consume(const CtxF& c, const PipeF& f, Row& cur) {
        const Geo& g = c.g;
        const int lane = g.lane;
        float Xc[S][C], lk[NL4 * 4];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float fx = f.fx[s], fy = f.fy[s];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                const float v00 = f.G[s][ch][0], v01 = f.G[s][ch][1], v10 = f.G[s][ch][2], v11 = f.G[s][ch][3];
                const float dtop = v01 - v00, dbot = v11 - v10, dl = v10 - v00;
                const float dd = dbot - dtop;
                const float dix = fmaf(fy, dd, dtop);          // d value / d ix
                const float diy = fmaf(fx, dd, dl);            // d value / d iy
                Xc[s][ch] = fmaf(fy, diy, fmaf(fx, dtop, v00)) - c.rc[ch];
                if (BWD) {
                    lk[(O_XM - O_XM) + s * C + ch] = Xc[s][ch];
                    lk[(O_SL - O_XM) + s * 2 * C + ch] = dix; lk[(O_SL - O_XM) + s * 2 * C + C + ch] = diy;
                }
            }
        }
        if (BWD) {
#pragma unroll
            for (int k = NLF; k < NL4 * 4; ++k) lk[k] = 0.f;
            const int base = slot_vec(c.fill) + NE4 * 32;
#pragma unroll
            for (int k = 0; k < NL4; ++k) {
                Vec4 s4; s4.x = lk[4 * k]; s4.y = lk[4 * k + 1]; s4.z = lk[4 * k + 2]; s4.w = lk[4 * k + 3];
                s_st4(c.ring, base + k * 32, s4);
            }
        }
        cur.D = f.d;
        // horizontal 3-sums (window column centred on this lane)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            const float Tc = f.Tc[ch];
            const float yl = w_up(Tc, lane), yr = w_dn(Tc, lane);
            cur.ym[ch] = Tc;
            cur.hy[ch] = yl + Tc + yr;
            cur.hyy[ch] = fmaf(yr, yr, fmaf(Tc, Tc, yl * yl));
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float xl = w_up(Xc[s][ch], lane), xr = w_dn(Xc[s][ch], lane);
                cur.xm[s][ch] = Xc[s][ch];
                cur.hx[s][ch] = xl + Xc[s][ch] + xr;
                cur.hxx[s][ch] = fmaf(xr, xr, fmaf(Xc[s][ch], Xc[s][ch], xl * xl));
                cur.hxy[s][ch] = fmaf(xr, yr, fmaf(Xc[s][ch], Tc, xl * yl));
            }
        }
    }

    // ---- W(i-1): windows centred on row q = i-1; rows a = i-2, b = i-1, cur = i ----
    static MD2_DEV void stage_windows(const FusedParams& p, const CtxF& c, AccF& acc, const Row& a, const Row& b,
                                      const Row& cur, int i, float (&wp)[NWP4 * 4]) {
        const Geo& g = c.g;
        const int lane = g.lane;
        const int q = i - 1;
        const bool row_in = q >= 0 && q < g.H;
        const bool row_own = q >= g.Y0 && q < g.Y1;
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
        if (c.has_am) {
            MD2_RARE_BLOCK();
            const int qc = q < 0 ? 0 : (q > g.H - 1 ? g.H - 1 : q);
            const float am = g_ld(c.am + (qc * g.W + g.gxm));
            if (am <= wlv) { wlv = am; sel = -1; }   // mask is first in the cat: wins ties
        }
        const bool own = row_own && g.pcol;
        acc.warp_sum += own ? wlv : 0.f;
        if (c.do_viz) {
            MD2_RARE_BLOCK();
            const long long o = (long long)g.n * g.HW + q * g.W + g.gxm;   // (gxm == gxr on the output columns)
            if (p.viz_loss && own) p.viz_loss[o] = wlv;
#pragma unroll
            for (int s = 0; s < S; ++s)
                if (p.viz_warped[s] && own) {
#pragma unroll
                    for (int ch = 0; ch < C; ++ch)
                        p.viz_warped[s][((long long)g.n * C + ch) * g.HW + q * g.W + g.gxm] = b.xm[s][ch] + c.rc[ch];
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
                if (g.has_right) acc.ssx += fabsf(b.D - Dr) * MD2_EXP(-gx_ * (1.0f / C));
                if (q + 1 < g.H) acc.ssy += fabsf(b.D - cur.D) * MD2_EXP(-gy_ * (1.0f / C));
                acc.dsum += b.D;
            }
        }
        if (BWD) {
            // scale by the upstream cotangent of this window (0 outside the image / where the automask won);
            // warp B forms the horizontal adjoint sums
            const float k = (row_in && sel >= 0) ? c.kq : 0.f;
#pragma unroll
            for (int j = 0; j < 3 * C; ++j) wp[j] = cf[j] * k;
            wp[3 * C] = i_as_float(sel);
            // smoothness gradient of pixel row q before the mean-normalisation (warp B applies A ghat - B):
            // ghat = (ex(q) - ex(q)[left lane]) + (ey(q) - ey(q-1)), e = sign(d - d') exp(-mean_c |T - T'|) / count
            if (MD2_SMOOTH_IN_B) wp[3 * C + 1] = 0.f;
            else {
                const float Dr = w_dn(b.D, lane);
                float gxs = 0.f, gys = 0.f;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    gxs += fabsf(b.ym[ch] - w_dn(b.ym[ch], lane));
                    gys += fabsf(b.ym[ch] - cur.ym[ch]);
                }
                const float wx = f_ex2(gxs * (-1.4426950408889634f / C)) * c.cxn;
                const float wy = f_ex2(gys * (-1.4426950408889634f / C)) * c.cyn;
                const float ex = g.has_right ? sgn_scaled(b.D - Dr, wx) : 0.f;
                const float ey = (q >= 0 && q + 1 < g.H) ? sgn_scaled(b.D - cur.D, wy) : 0.f;
                const float exl = w_up(ex, lane);
                wp[3 * C + 1] = (ex - exl) + (ey - acc.ey_prev);
                acc.ey_prev = ey;
            }
#pragma unroll
            for (int k2 = NWPF; k2 < NWP4 * 4; ++k2) wp[k2] = 0.f;
        }
    }

    // the window packet completes the slot of the consumed row: store it and hand the slot to warp B
    static MD2_DEV void finish_slot(CtxF& c, const float (&wp)[NWP4 * 4]) {
        const int base = slot_vec(c.fill) + NPP4 * 32;
#pragma unroll
        for (int k = 0; k < NWP4; ++k) {
            Vec4 s4; s4.x = wp[4 * k]; s4.y = wp[4 * k + 1]; s4.z = wp[4 * k + 2]; s4.w = wp[4 * k + 3];
            s_st4(c.ring, base + k * 32, s4);
        }
        mb_arrive(c.bars, BAR_FULL + c.fill.idx);
        c.fill = next_slot(c.fill);
    }

    // one row of the pipeline: C(i), A(i+1), W(i-1)
    static MD2_DEV void stepF(const FusedParams& p, CtxF& c, PipeF& f, AccF& acc, Row& a, Row& b, Row& cur, int i, int iend,
                              bool windows) {
        float wp[NWP4 * 4];
        stage_consume(c, f, cur);
        if (i + 1 < iend) stage_issue(p, c, f, i + 1, next_slot(c.fill));
        if (windows) stage_windows(p, c, acc, a, b, cur, i, wp);
        else {
#pragma unroll
            for (int k = 0; k < NWP4 * 4; ++k) wp[k] = 0.f;
        }
        if (BWD) finish_slot(c, wp);
    }

    // warp F of one work item; gslot: running ring-slot counter of this warp
    static MD2_DEV void run_forward(const FusedParams& p, int sx, int cy, int z, int lane, float* wsm, int& gslot, float (&v)[32]) {
        CtxF c;
        geometry(p, sx, cy, z, lane, c.g);
        const Geo& g = c.g;
        const bool wcol = g.col_img && lane >= 1 && lane <= 30;             // window column
        c.Wf = (float)g.W; c.Hf = (float)g.H;
        c.do_viz = g.scale == p.L - 1 && (p.viz_loss != nullptr || p.viz_warped[0] != nullptr || (S > 1 && p.viz_warped[S - 1] != nullptr));
        const float* tgn = p.tgt + (long long)g.n * p.tgt_ns;
        c.tg = tgn + g.gxm;
#pragma unroll
        for (int s = 0; s < S; ++s) c.sb[s] = p.src[s] + (long long)g.n * p.src_ns[s];
        c.dp = p.dfull[g.scale] + (long long)g.n * g.HW + g.gxm;
        c.am = p.automask ? p.automask + (long long)g.n * g.HW : nullptr;
        c.has_am = p.automask != nullptr;
        // centring constant of the window sums (any constant is exact; a local value keeps the
        // centred squares small): the target at the middle of the strip chunk
        {
            const int ym = (g.Y0 + g.Y1) >> 1;
            const int xm = sx * OW + OW / 2 < g.W ? sx * OW + OW / 2 : g.W - 1;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) c.rc[ch] = tgn[ch * g.HW + ym * g.W + xm];
        }
        const float px = (float)(g.gxm + 1);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            c.pb[s] = p.pose_slot + (s * p.N + g.n) * 12;
#pragma unroll
            for (int k = 0; k < 3; ++k)   // A[:,0] px + A[:,2]: the lane-constant part of A p
                c.apx[s][k] = fmaf(MD2_POSE(p, c.pb[s] + 3 * k), px, MD2_POSE(p, c.pb[s] + 3 * k + 2));
        }
        const float up_photo = p.gloss * p.loss_scale / ((float)g.W * (float)g.H * (float)p.N);
        c.kq = wcol ? up_photo * (PHOTO_ALPHA / C) * (-0.5f) : 0.f;
        c.cxn = 1.0f / ((float)(g.W - 1) * (float)g.H * (float)p.N);
        c.cyn = 1.0f / ((float)g.W * (float)(g.H - 1) * (float)p.N);
        c.ring = ring_ref_of(wsm, lane);
        c.bars = bar_ref_of(wsm + RING_FLOATS);
        c.fill = slot_of(gslot);
        // pin the per-lane invariants in registers (otherwise they are re-derived in every row)
        keep(c.g.gxm); if (PIN_MORE) { keep(c.g.lane); keep(c.g.W); keep(c.has_am); } keep(c.tg); keep(c.dp); keep(c.ring); keep(c.bars);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            keep(c.sb[s]);
#pragma unroll
            for (int k = 0; k < 3; ++k) keep(c.apx[s][k]);
        }
        if (BWD) keep(c.kq);
#pragma unroll
        for (int ch = 0; ch < C; ++ch) keep(c.rc[ch]);

        AccF acc;
        acc.warp_sum = acc.ssx = acc.ssy = acc.dsum = acc.ey_prev = 0.f;
        Row r0, r1, r2;
        PipeF f;
        // rows i0, i0+1 carry pixel packets only (no window yet); then every row runs C(i), A(i+1), W(i-1);
        // the loop is unrolled by 3 so that the three row registers rotate roles without moves
        const int i0 = g.Y0 - HALO, iend = g.Y1 + HALO;
        {   // prime the pipeline: loads of row i0, then A(i0)
            f.gy = image_row(i0, g.H);
            const int toff = f.gy * g.W;
            f.dn = g_ld(c.dp + toff);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) f.Tn[ch] = g_ld(c.tg + (ch * g.HW + toff));
            stage_issue(p, c, f, i0, c.fill);
        }
        stepF(p, c, f, acc, r0, r0, r0, i0, iend, false);
        stepF(p, c, f, acc, r0, r0, r1, i0 + 1, iend, false);
        int i = i0 + 2;
        for (; i + 2 < iend; i += 3) {
            stepF(p, c, f, acc, r0, r1, r2, i, iend, true);
            stepF(p, c, f, acc, r1, r2, r0, i + 1, iend, true);
            stepF(p, c, f, acc, r2, r0, r1, i + 2, iend, true);
        }
        if (i < iend) {
            stepF(p, c, f, acc, r0, r1, r2, i, iend, true);
            if (i + 1 < iend) stepF(p, c, f, acc, r1, r2, r0, i + 1, iend, true);
        }
        gslot = counter_of(c.fill);
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;
        v[0] = acc.warp_sum; v[1] = acc.ssx; v[2] = acc.ssy; v[3] = acc.dsum;
    }

    // =====================================================================================
    // warp B
    // =====================================================================================
    struct CtxB {
        Geo g;
        float* gb[S];
        int has_gb[S];               // (kept in registers, like has_am)
        float* gd;                   // full-resolution disparity gradient of this (scale, image)
        float apx[S][3];
        int pb[S];
        float cl1;                   // up_photo * (1-alpha)/C
        float mp;                    // pcol ? 1 : 0
        float cxn, cyn, sA, sB, nega;
        float wl, wr;                // horizontal reflect-pad adjoint weights of this pixel column
        float Wf, Hf;
        ring_ref ring;               // + this lane's Vec4 column
        bar_ref bars;
    };
    // one window row after the horizontal adjoint 3-sums: t = sums of the (scaled) coefficients of the
    // three windows around this column, z = the part of t selected for source 0, sel = this column's selection
    struct WinRow { float t[3 * C], z[3 * C]; int sel; float gh; };
    struct AccB {
        float ey_prev;               // vertical smoothness edge between the previous row and this one (MD2_SMOOTH_IN_B)
        float P0[S][3], P1[S][3], Ph[S][3];
        float car0[S][C], car1[S][C];
        int coff[S];
    };

    // acquire = wait until warp F has filled the slot, release = hand it back
    static MD2_DEV void acquire(const CtxB& c, Slot t) { mb_wait(c.bars, BAR_FULL + t.idx, t.par); }
    static MD2_DEV void release(const CtxB& c, Slot t) { mb_arrive(c.bars, BAR_EMPTY + t.idx); }

    // window row carried by slot `t` (window packet of the row above the slot's pixel row): horizontal adjoint 3-sums
    static MD2_DEV void load_window_row(const CtxB& c, Slot t, WinRow& w) {
        const int lane = c.g.lane;
        const int sv = slot_vec(t) + NPP4 * 32;
        float wq[NWP4 * 4];
#pragma unroll
        for (int k = 0; k < NWP4; ++k) {
            const Vec4 q = s_ld4(c.ring, sv + k * 32);
            wq[4 * k] = q.x; wq[4 * k + 1] = q.y; wq[4 * k + 2] = q.z; wq[4 * k + 3] = q.w;
        }
        const int sel = f_as_int(wq[3 * C]);
        w.sel = sel;
        w.gh = wq[3 * C + 1];
        const int e0 = w_up(sel, lane), e2 = w_dn(sel, lane);
        const float m0 = (e0 == 0) ? c.wl : 0.f, m1 = (sel == 0) ? 1.f : 0.f, m2 = (e2 == 0) ? c.wr : 0.f;
#pragma unroll
        for (int j = 0; j < 3 * C; ++j) {
            const float c1 = wq[j];
            const float c0 = w_up(c1, lane), c2 = w_dn(c1, lane);
            w.t[j] = fmaf(c.wl, c0, fmaf(c.wr, c2, c1));
            w.z[j] = (S > 1) ? fmaf(m0, c0, fmaf(m2, c2, m1 * c1)) : 0.f;
        }
    }

    // ---- P(r): slot t0 = row r (pixel packet); wa, wb, wc = window rows r-1, r, r+1 ----
    // vertical smoothness edge between rows y and y+1 (own target values / disparities of both rows)
    static MD2_DEV float edge_y(const CtxB& c, int y, const float* ymA, float DA, const float* ymB, float DB) {
        float gsum = 0.f;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) gsum += fabsf(ymA[ch] - ymB[ch]);
        const float w = f_ex2(gsum * (-1.4426950408889634f / C)) * c.cyn;
        return (y >= 0 && y + 1 < c.g.H) ? sgn_scaled(DA - DB, w) : 0.f;
    }
    // ym[C], D of the row in slot t
    static MD2_DEV void load_ym_d(const CtxB& c, Slot t, float (&yd)[NYD4 * 4]) {
        const int sv = slot_vec(t);
#pragma unroll
        for (int k = 0; k < NYD4; ++k) {
            const Vec4 q = s_ld4(c.ring, sv + k * 32);
            yd[4 * k] = q.x; yd[4 * k + 1] = q.y; yd[4 * k + 2] = q.z; yd[4 * k + 3] = q.w;
        }
    }

    static MD2_DEV void stage_pixels(const FusedParams& p, const CtxB& c, AccB& acc, int r, Slot t0, Slot t1, const WinRow& wa,
                                     const WinRow& wb, const WinRow& wc) {
        const Geo& g = c.g;
        const int lane = g.lane;
        const int s0 = slot_vec(t0);
        float pk[NPP4 * 4];
#pragma unroll
        for (int k = 0; k < NPP4; ++k) {
            const Vec4 t = s_ld4(c.ring, s0 + k * 32);
            pk[4 * k] = t.x; pk[4 * k + 1] = t.y; pk[4 * k + 2] = t.z; pk[4 * k + 3] = t.w;
        }
        const float wu = (r == 1) ? 2.f : 1.f, wd = (r == g.H - 2) ? 2.f : 1.f;
        const float pyr = (float)(r + 1);
        const int selr = wb.sel;
        const float zr = pk[O_Z];
        float dbar_z = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float* st = pk + O_SRC + s * SRCF;
            // d loss / d warped_s at this pixel
            float ibar[C];
            bool act = false;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                float sa, sb_, sg;
                const float ta = fmaf(wu, wa.t[3 * ch], fmaf(wd, wc.t[3 * ch], wb.t[3 * ch]));
                const float tb = fmaf(wu, wa.t[3 * ch + 1], fmaf(wd, wc.t[3 * ch + 1], wb.t[3 * ch + 1]));
                const float tgm = fmaf(wu, wa.t[3 * ch + 2], fmaf(wd, wc.t[3 * ch + 2], wb.t[3 * ch + 2]));
                if (S == 1) { sa = ta; sb_ = tb; sg = tgm; }
                else {
                    const float za = fmaf(wu, wa.z[3 * ch], fmaf(wd, wc.z[3 * ch], wb.z[3 * ch]));
                    const float zb = fmaf(wu, wa.z[3 * ch + 1], fmaf(wd, wc.z[3 * ch + 1], wb.z[3 * ch + 1]));
                    const float zg = fmaf(wu, wa.z[3 * ch + 2], fmaf(wd, wc.z[3 * ch + 2], wb.z[3 * ch + 2]));
                    if (s == 0) { sa = za; sb_ = zb; sg = zg; }
                    else { sa = ta - za; sb_ = tb - zb; sg = tgm - zg; }
                }
                const float xj = pk[O_XM + s * C + ch], yj = pk[O_YM + ch];
                float gv = fmaf(xj, sb_, fmaf(yj, sg, sa));
                if (selr == s) gv += sgn_scaled(xj - yj, c.cl1);
                ibar[ch] = gv * c.mp;
                act = act || (ibar[ch] != 0.f);
            }
            // sources not selected anywhere in the 3x3 neighbourhood of any lane skip all of this
            if (!MD2_SKIP_IDLE_SOURCE || w_any(act)) {
                const float q = st[0], u = st[1], vv = st[2];
                const float fx = st[3], fy = st[4];
                const int off = f_as_int(st[5]);
                // clip-gradient masks (0 where the un-clipped coordinate is <= 1 or >= size), folded into q
                const float qa = (u > 1.0f && u < c.Wf) ? q : 0.0f;
                const float qb = (vv > 1.0f && vv < c.Hf) ? q : 0.0f;
                float du = 0.f, dv = 0.f;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    du = fmaf(ibar[ch], pk[O_SL + s * 2 * C + ch], du);
                    dv = fmaf(ibar[ch], pk[O_SL + s * 2 * C + C + ch], dv);
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
#if MD2_SCATTER_CARRY
                // source-image gradient: scatter with vertical carry + merge with the right-hand lane
                if (c.has_gb[s]) {
                    const bool sval = act;     // (ibar is already 0 outside the output columns)
                    const float gx1 = 1.f - fx, gy1 = 1.f - fy;
                    float tq0[C], tq1[C];
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) { tq0[ch] = gx1 * gy1 * ibar[ch]; tq1[ch] = fx * gy1 * ibar[ch]; }
                    const int key = sval ? off : -2;
                    const bool have = acc.coff[s] >= 0;
                    const bool aligned = have && key == acc.coff[s] + g.W;
                    if (aligned) {
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) { tq0[ch] += acc.car0[s][ch]; tq1[ch] += acc.car1[s][ch]; }
                    } else if (have) {
                        float* o = c.gb[s] + (acc.coff[s] + g.W);
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) {
                            g_red(o + ch * g.HW, acc.car0[s][ch]);
                            g_red1(o + ch * g.HW, acc.car1[s][ch]);
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
                            g_red(o + ch * g.HW, tq0[ch]);
                            if (!absorbed) g_red1(o + ch * g.HW, tq1[ch]);
                            acc.car0[s][ch] = gx1 * fy * ibar[ch];
                            acc.car1[s][ch] = fx * fy * ibar[ch];
                        }
                    }
                    acc.coff[s] = sval ? off : -1;
                }
#else
                // source-image gradient: both tap rows now, right taps merged into the right-hand lane's left taps
                if (c.has_gb[s]) {
                    const bool sval = act;     // (ibar is already 0 outside the output columns)
                    const float gx1 = 1.f - fx, gy1 = 1.f - fy;
#if MD2_SCATTER_MERGE
                    const int key = sval ? off : -2;
                    const int key_r = w_dn(key, lane), key_l = w_up(key, lane);
                    const bool absorbed = sval && lane < 31 && key_r == key + 1;
                    const bool absorb = sval && lane > 0 && key_l >= 0 && key_l + 1 == key;
#else
                    const bool absorbed = false;
#endif
                    float* o = c.gb[s] + off;
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) {
                        const float wl = gx1 * ibar[ch], wr = fx * ibar[ch];
                        float t0 = wl * gy1, t1 = wr * gy1, b0 = wl * fy, b1 = wr * fy;
#if MD2_SCATTER_MERGE
                        const float flt = w_up(t1, lane), flb = w_up(b1, lane);
                        if (absorb) { t0 += flt; b0 += flb; }
#endif
                        if (sval) {
                            g_red(o + ch * g.HW, t0);
                            g_red(o + (ch * g.HW + g.W), b0);
                            if (!absorbed) { g_red1(o + ch * g.HW, t1); g_red1(o + (ch * g.HW + g.W), b1); }
                        }
                    }
                }
#endif
            }
#if MD2_SCATTER_CARRY
            else if (c.has_gb[s]) {
                // nobody scatters into source s on this row: flush what the previous row carried
                if (acc.coff[s] >= 0) {
                    float* o = c.gb[s] + (acc.coff[s] + g.W);
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) {
                        g_red(o + ch * g.HW, acc.car0[s][ch]);
                        g_red1(o + ch * g.HW, acc.car1[s][ch]);
                    }
                }
                acc.coff[s] = -1;
            }
#endif
        }
        // depth -> disparity:  dz/dd = -a z^2
        float gd = c.nega * zr * zr * dbar_z;
        // smoothness gradient (src/utils.jl:159-173 with the mean-normalisation of
        // src/training.jl:64-65 folded in):  A ghat_j - B   (ghat comes from warp F with the window row)
        float gh = wb.gh;
        if (MD2_SMOOTH_IN_B) {
            float nx[NYD4 * 4];
            load_ym_d(c, t1, nx);
            const float Da = pk[O_D];
            const float ey = edge_y(c, r, pk + O_YM, Da, nx + O_YM, nx[O_D]);
            const float Dr = w_dn(Da, lane);
            float gsum = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) gsum += fabsf(pk[O_YM + ch] - w_dn(pk[O_YM + ch], lane));
            const float w = f_ex2(gsum * (-1.4426950408889634f / C)) * c.cxn;
            const float ex = g.has_right ? sgn_scaled(Da - Dr, w) : 0.f;
            const float exl = w_up(ex, lane);
            gh = (ex - exl) + (ey - acc.ey_prev);
            acc.ey_prev = ey;
        }
        gd += fmaf(c.sA, gh, -c.sB);
        if (g.pcol) g_st(c.gd + (r * g.W + g.gxm), gd);   // (gxm == gxr on the output columns)
    }

    // warp B of one work item; gslot: running ring-slot counter of this warp
    static MD2_DEV void run_backward(const FusedParams& p, int sx, int cy, int z, int lane, float* wsm, int& gslot, float (&v)[32]) {
        CtxB c;
        geometry(p, sx, cy, z, lane, c.g);
        const Geo& g = c.g;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            c.gb[s] = p.gsrc[s] ? p.gsrc[s] + (long long)g.n * p.src_ns[s] : nullptr;
            c.has_gb[s] = p.gsrc[s] != nullptr;
            c.pb[s] = p.pose_slot + (s * p.N + g.n) * 12;
        }
        c.gd = p.gfull[g.scale] + (long long)g.n * g.HW;
        const float px = (float)(g.gxm + 1);
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                c.apx[s][k] = fmaf(MD2_POSE(p, c.pb[s] + 3 * k), px, MD2_POSE(p, c.pb[s] + 3 * k + 2));
        const float up_photo = p.gloss * p.loss_scale / ((float)g.W * (float)g.H * (float)p.N);
        c.cl1 = up_photo * ((1.0f - PHOTO_ALPHA) / C);
        c.mp = g.pcol ? 1.f : 0.f;
        c.cxn = 1.0f / ((float)(g.W - 1) * (float)g.H * (float)p.N);
        c.cyn = 1.0f / ((float)g.W * (float)(g.H - 1) * (float)p.N);
        c.nega = -p.depth_a;
        c.wl = (g.gxr == 1) ? 2.f : 1.f;
        c.wr = (g.gxr == g.W - 2) ? 2.f : 1.f;
        c.Wf = (float)g.W; c.Hf = (float)g.H;
        {
            const float up_s = p.gloss * p.loss_scale * p.smooth_w[g.scale];
            c.sA = up_s; c.sB = 0.f;
            if (p.normalize_disp) {
                float ssx, ssy, dsum;
                prep_stats(p, g.scale, g.n, lane, ssx, ssy, dsum);
                const float m = dsum / (float)g.HW + 1e-7f;
                c.sA = up_s / m;
                c.sB = up_s * (c.cxn * ssx + c.cyn * ssy) / (m * m * (float)g.HW);
            }
        }
        c.ring = ring_ref_of(wsm, lane);
        c.bars = bar_ref_of(wsm + RING_FLOATS);
        keep(c.g.gxm); if (PIN_MORE) { keep(c.g.lane); keep(c.g.W); } keep(c.gd); keep(c.cl1); keep(c.mp); keep(c.sA); keep(c.sB); keep(c.ring); keep(c.bars);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            keep(c.gb[s]); if (PIN_MORE) keep(c.has_gb[s]);
#pragma unroll
            for (int k = 0; k < 3; ++k) keep(c.apx[s][k]);
        }
        AccB acc;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            acc.coff[s] = -1;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc.P0[s][k] = acc.P1[s][k] = acc.Ph[s][k] = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) acc.car0[s][ch] = acc.car1[s][ch] = 0.f;
        }
        // the ring slots after `gslot` hold rows Y0-2, Y0-1, Y0, ... of this item
        Slot ta = slot_of(gslot);
        acquire(c, ta);                                            // row Y0-2: nothing to read
        release(c, ta);
        ta = next_slot(ta);
        acquire(c, ta);                                            // row Y0-1
        Slot tb = next_slot(ta);
        acquire(c, tb);                                            // row Y0
        acc.ey_prev = 0.f;
        if (MD2_SMOOTH_IN_B) {   // the "up" edge of the first row
            float ya[NYD4 * 4], yb[NYD4 * 4];
            load_ym_d(c, ta, ya);
            load_ym_d(c, tb, yb);
            acc.ey_prev = edge_y(c, g.Y0 - 1, ya + O_YM, ya[O_D], yb + O_YM, yb[O_D]);
        }
        release(c, ta);
        ta = tb;                                                   // ta = slot of row r, tb = slot of row r+1
        tb = next_slot(tb);
        acquire(c, tb);                                            // row Y0+1
        // slot(i) carries window row i-1; the three window rows around pixel row r rotate through w0, w1, w2
        // (loop unrolled by 3: no register moves)
        WinRow w0, w1, w2;
        load_window_row(c, ta, w0);                                // window row Y0-1
        load_window_row(c, tb, w1);                                // window row Y0
        int r = g.Y0;
#define MD2_STEP_B(WA, WB, WC)                                                                        \
        {                                                                                             \
            const Slot tc = next_slot(tb);                                                            \
            acquire(c, tc);                                        /* row r+2: carries window row r+1 */ \
            load_window_row(c, tc, WC);                                                               \
            stage_pixels(p, c, acc, r, ta, tb, WA, WB, WC);                                           \
            release(c, ta);                                                                           \
            ta = tb; tb = tc;                                                                         \
            ++r;                                                                                      \
        }
#if MD2_B_UNROLL3
        for (; r + 2 < g.Y1;) {
            MD2_STEP_B(w0, w1, w2)
            MD2_STEP_B(w1, w2, w0)
            MD2_STEP_B(w2, w0, w1)
        }
        if (r < g.Y1) {
            MD2_STEP_B(w0, w1, w2)
            if (r < g.Y1) MD2_STEP_B(w1, w2, w0)
        }
#else
        while (r < g.Y1) {
            MD2_STEP_B(w0, w1, w2)
            w0 = w1; w1 = w2;
        }
#endif
#undef MD2_STEP_B
        release(c, ta);                                            // rows Y1, Y1+1
        release(c, tb);
        gslot = counter_of(next_slot(tb));
        // flush the carried lower tap pairs of the last row
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (MD2_SCATTER_CARRY && c.gb[s] && acc.coff[s] >= 0) {
                float* o = c.gb[s] + (acc.coff[s] + g.W);
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    g_red(o + ch * g.HW, acc.car0[s][ch]);
                    g_red1(o + ch * g.HW, acc.car1[s][ch]);
                }
            }
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int k = 0; k < 3; ++k) {   // G = sum cbar (z p)^T, p = (px, py, 1); h = sum cbar
                v[NSTAT + 12 * s + 3 * k + 0] = px * acc.P0[s][k];
                v[NSTAT + 12 * s + 3 * k + 1] = acc.P1[s][k];
                v[NSTAT + 12 * s + 3 * k + 2] = acc.P0[s][k];
                v[NSTAT + 12 * s + 9 + k] = acc.Ph[s][k];
            }
    }
};

}  // namespace md2
