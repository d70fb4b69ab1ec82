// Fused view-synthesis loss: __global__ wrappers around md2_fused.cuh + host orchestration
// + the C ABI (md2_view_synthesis_loss_{fwd,bwd,fwdbwd}).  sm_100a only.
#include <stdarg.h>
#include <string.h>

#include "md2_common.cuh"
#include "md2_fused.cuh"

namespace md2 {

// ------------------------------------------------------------------------------------------
// error / ctx plumbing
// ------------------------------------------------------------------------------------------
std::string& last_error_ref() {
    static thread_local std::string e;
    return e;
}
int set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return 1;
}
void* ws_get(md2_ctx* ctx, int slot, size_t bytes) {
    Workspace& w = ctx->ws[slot];
    if (w.bytes >= bytes && w.ptr) return w.ptr;
    if (w.ptr) {
        cudaDeviceSynchronize();  // growth only: a previous launch may still read the old buffer
        cudaFree(w.ptr);
        w.ptr = nullptr; w.bytes = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&w.ptr, want);
    if (e != cudaSuccess) {
        set_error("workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        w.ptr = nullptr;
        return nullptr;
    }
    w.bytes = want;
    return w.ptr;
}

// ------------------------------------------------------------------------------------------
// pose preparation: [composeT] + pre-composition  A = K R K^-1, b = K t      (A3, A4, A6)
// ------------------------------------------------------------------------------------------
struct PoseArgs {
    int S, N, mode;
    const float* rot[MAX_S];
    const float* trans[MAX_S];
    int invert[MAX_S];
    const float* K; const float* invK;
    float* grot[MAX_S];
    float* gtrans[MAX_S];
};

__global__ void pose_prep_kernel(PoseArgs a, float* __restrict__ pose_ab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.S * a.N) return;
    const int s = i / a.N, n = i % a.N;
    double K[9], Ki[9], R[9], t[3];
    load_cm3(a.K, K);
    load_cm3(a.invK, Ki);
    if (a.mode == 0) {
        load_cm3(a.rot[s] + 9 * n, R);
        for (int k = 0; k < 3; ++k) t[k] = a.trans[s][3 * n + k];
    } else {
        double r[3], tv[3];
        for (int k = 0; k < 3; ++k) { r[k] = a.rot[s][3 * n + k]; tv[k] = a.trans[s][3 * n + k]; }
        compose_T(r, tv, a.invert[s], R, t);
    }
    precompose(K, Ki, R, t, pose_ab + ((long long)s * a.N + n) * 12);
}

// ------------------------------------------------------------------------------------------
// align-corners bilinear upsample of the low-resolution disparities (A17) and its adjoint
// ------------------------------------------------------------------------------------------

__global__ void upsample_kernel(const float* __restrict__ in, float* __restrict__ out, int w, int h,
                                int W, int H, int CN) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)W * H * CN) return;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long cn = i / ((long long)W * H);
    const float sx = up_scale(w, W), sy = up_scale(h, H);
    int x0, x1, y0, y1; float fx, fy;
    up_taps(x, sx, w, x0, x1, fx);
    up_taps(y, sy, h, y0, y1, fy);
    const float* b = in + cn * w * h;
    out[i] = bilerp(b[y0 * w + x0], b[y0 * w + x1], b[y1 * w + x0], b[y1 * w + x1], fx, fy);
}

// gather form (deterministic): one thread per low-resolution pixel
__global__ void upsample_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int w,
                                    int h, int W, int H, int CN) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)w * h * CN) return;
    const int xi = (int)(i % w), yi = (int)((i / w) % h);
    const long long cn = i / ((long long)w * h);
    const float sx = up_scale(w, W), sy = up_scale(h, H);
    // candidate full-res range whose taps can touch (xi, yi)
    int xlo = 0, xhi = W - 1, ylo = 0, yhi = H - 1;
    if (sx > 0.f) {
        xlo = max(0, (int)floorf((float)(xi - 1) / sx) - 1);
        xhi = min(W - 1, (int)ceilf((float)(xi + 1) / sx) + 1);
    }
    if (sy > 0.f) {
        ylo = max(0, (int)floorf((float)(yi - 1) / sy) - 1);
        yhi = min(H - 1, (int)ceilf((float)(yi + 1) / sy) + 1);
    }
    const float* g = gout + cn * W * H;
    float acc = 0.f;
    for (int y = ylo; y <= yhi; ++y) {
        int y0, y1; float fy;
        up_taps(y, sy, h, y0, y1, fy);
        const float wy = (y0 == yi ? 1.f - fy : 0.f) + (y1 == yi ? fy : 0.f);
        if (wy == 0.f) continue;
        float row = 0.f;
        for (int x = xlo; x <= xhi; ++x) {
            int x0, x1; float fx;
            up_taps(x, sx, w, x0, x1, fx);
            const float wx = (x0 == xi ? 1.f - fx : 0.f) + (x1 == xi ? fx : 0.f);
            row = fmaf(wx, g[y * W + x], row);
        }
        acc = fmaf(wy, row, acc);
    }
    gin[i] = acc;
}

// ------------------------------------------------------------------------------------------
// smoothness / mean-disparity statistics pre-pass (needed before the fused backward because
// d / mean(d) couples every pixel of an image)
// ------------------------------------------------------------------------------------------
struct StatsArgs {
    int W, H, N, L;
    const float* tgt; long long tgt_ns;
    const float* disp[MAX_L];
};

template <int C>
__global__ void __launch_bounds__(256) stats_kernel(StatsArgs a, float* __restrict__ partial) {
    __shared__ float scratch[3 * 8];
    const int z = blockIdx.y, scale = z / a.N, n = z % a.N;
    const long long HW = (long long)a.W * a.H;
    const float* d = a.disp[scale] + n * HW;
    const float* t = a.tgt + n * a.tgt_ns;
    float v[3] = {0.f, 0.f, 0.f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW;
         i += (long long)gridDim.x * blockDim.x) {
        stats_pixel<C>(d, t, i, a.W, a.H, v[0], v[1], v[2]);
    }
    block_sum<3>(v, scratch);
    if (threadIdx.x == 0) {
        float* o = partial + ((long long)z * gridDim.x + blockIdx.x) * NSTAT;
        o[0] = 0.f; o[1] = v[0]; o[2] = v[1]; o[3] = v[2];
    }
}

// out[g][k] = sum_b partial[g][b][k], fixed order (deterministic).  NP <= 32.
// Columns k < n0 go to out0 (row length n0), the rest to out1 (row length NP - n0).
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial,
                                                              float* __restrict__ out0, int n0,
                                                              float* __restrict__ out1, int bpg, int NP) {
    __shared__ float sm[8][32];
    const int g = blockIdx.x, k = threadIdx.x & 31, chunk = threadIdx.x >> 5;
    float s = 0.f;
    if (k < NP)
        for (int b = chunk; b < bpg; b += 8) s += partial[((long long)g * bpg + b) * NP + k];
    sm[chunk][k] = s;
    __syncthreads();
    if (chunk == 0 && k < NP) {
        float t = 0.f;
        for (int c = 0; c < 8; ++c) t += sm[c][k];
        if (k < n0) out0[(long long)g * n0 + k] = t;
        else if (out1) out1[(long long)g * (NP - n0) + (k - n0)] = t;
    }
}

int launch_reduce_partials(md2_ctx* ctx, const float* partial, float* out0, int n0, float* out1,
                           int groups, int bpg, int NP, cudaStream_t st) {
    reduce_partials_kernel<<<groups, 256, 0, st>>>(partial, out0, n0, out1, bpg, NP);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------
// the fused tile kernel
// ------------------------------------------------------------------------------------------
template <int C, int S, bool BWD>
__global__ void __launch_bounds__(FUSED_THREADS) fused_kernel(const __grid_constant__ FusedParams p) {
    extern __shared__ float sm[];
    using F = Fused<C, S, BWD>;
    const int z = blockIdx.z, scale = z / p.N, n = z % p.N;
    const int tx0 = blockIdx.x * TILE_W, ty0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.x;
    FusedAcc<S> acc;
    acc.clear();
    F::phase_load(p, sm, scale, n, tx0, ty0, tid, FUSED_THREADS);
    __syncthreads();
    F::phase_windows(p, sm, scale, n, tx0, ty0, tid, FUSED_THREADS, acc);
    if (!BWD) {
        F::phase_smooth_fwd(p, sm, tx0, ty0, tid, FUSED_THREADS, acc);
    } else {
        __syncthreads();
        F::phase_pixel_bwd(p, sm, scale, n, tx0, ty0, tid, FUSED_THREADS, acc);
    }
    __syncthreads();   // shared memory is re-used as reduction scratch from here on
    constexpr int NP = F::NPART;
    float v[NP];
    v[0] = acc.warp_sum; v[1] = acc.sx; v[2] = acc.sy; v[3] = acc.dsum;
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int k = 0; k < 12; ++k) v[NSTAT + 12 * s + k] = acc.pose[s][k];
    constexpr int NW = FUSED_THREADS / 32;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        if (!BWD && k >= NSTAT) break;
        const float r = warp_sum(v[k]);
        if (lane == 0) sm[k * NW + wid] = r;
    }
    __syncthreads();
    if (tid < NP) {
        float r = 0.f;
        if (BWD || tid < NSTAT)
            for (int w = 0; w < NW; ++w) r += sm[tid * NW + w];
        const long long blk = ((long long)z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        p.partial[blk * NP + tid] = r;
    }
}

// ------------------------------------------------------------------------------------------
// finalize: loss scalar, saved statistics, pose gradients
// ------------------------------------------------------------------------------------------
struct FinalArgs {
    int W, H, N, L, S, NP;
    int mode;              // 0 fwd, 1 bwd, 2 fwdbwd
    const float* sums;     // (L*N, NP) from the fused kernel
    float* stats;          // (L*N, NSTAT)
    float* saved;          // nullable copy of stats for a later bwd
    float* loss;           // nullable
    float smooth_w[MAX_L];
    float loss_scale;
    int normalize_disp;
    PoseArgs pose;
};

__global__ void __launch_bounds__(128) finalize_kernel(FinalArgs a) {
    const int LN = a.L * a.N;
    if (a.mode != 1) {
        for (int z = threadIdx.x; z < LN; z += blockDim.x) {
            float* st = a.stats + (long long)z * NSTAT;
            const float* su = a.sums + (long long)z * a.NP;
            st[0] = su[0];
            if (a.mode == 0) { st[1] = su[1]; st[2] = su[2]; st[3] = su[3]; }
            if (a.saved)
                for (int k = 0; k < NSTAT; ++k) a.saved[(long long)z * NSTAT + k] = st[k];
        }
        __syncthreads();
        if (threadIdx.x == 0 && a.loss)
            *a.loss = loss_from_stats(a.stats, a.W, a.H, a.N, a.L, a.smooth_w, a.loss_scale, a.normalize_disp);
    }
    if (a.mode != 0) {
        double K[9], Ki[9];
        load_cm3(a.pose.K, K);
        load_cm3(a.pose.invK, Ki);
        for (int i = threadIdx.x; i < a.S * a.N; i += blockDim.x) {
            const int s = i / a.N, n = i % a.N;
            double G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, h[3] = {0, 0, 0};
            for (int l = 0; l < a.L; ++l) {
                const float* su = a.sums + ((long long)l * a.N + n) * a.NP + NSTAT + 12 * s;
                for (int k = 0; k < 9; ++k) G[k] += su[k];
                for (int k = 0; k < 3; ++k) h[k] += su[9 + k];
            }
            double Rub[9], tub[3];
            precompose_bwd(K, Ki, G, h, Rub, tub);
            if (a.pose.mode == 0) {
                if (a.pose.grot[s])
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) a.pose.grot[s][9 * n + 3 * c + r] = (float)Rub[3 * r + c];
                if (a.pose.gtrans[s])
                    for (int k = 0; k < 3; ++k) a.pose.gtrans[s][3 * n + k] = (float)tub[k];
            } else {
                double r[3], tv[3], rb[3], tb[3];
                for (int k = 0; k < 3; ++k) { r[k] = a.pose.rot[s][3 * n + k]; tv[k] = a.pose.trans[s][3 * n + k]; }
                compose_T_bwd(r, tv, a.pose.invert[s], Rub, tub, rb, tb);
                if (a.pose.grot[s])
                    for (int k = 0; k < 3; ++k) a.pose.grot[s][3 * n + k] = (float)rb[k];
                if (a.pose.gtrans[s])
                    for (int k = 0; k < 3; ++k) a.pose.gtrans[s][3 * n + k] = (float)tb[k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
template <int C, int S, bool BWD>
static int launch_fused(md2_ctx* ctx, const FusedParams& p, cudaStream_t st) {
    using F = Fused<C, S, BWD>;
    static bool configured = false;
    const size_t smem = sizeof(float) * (size_t)F::SMEM_FLOATS;
    if (!configured) {
        MD2_CHECK(cudaFuncSetAttribute(fused_kernel<C, S, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid(cdiv(p.W, TILE_W), cdiv(p.H, TILE_H), p.L * p.N);
    fused_kernel<C, S, BWD><<<grid, FUSED_THREADS, smem, st>>>(p);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

template <bool BWD>
static int dispatch_fused(md2_ctx* ctx, int C, int S, const FusedParams& p, cudaStream_t st) {
    if (C == 1 && S == 1) return launch_fused<1, 1, BWD>(ctx, p, st);
    if (C == 1 && S == 2) return launch_fused<1, 2, BWD>(ctx, p, st);
    if (C == 3 && S == 1) return launch_fused<3, 1, BWD>(ctx, p, st);
    if (C == 3 && S == 2) return launch_fused<3, 2, BWD>(ctx, p, st);
    return set_error("view_synthesis_loss: unsupported C=%d S=%d (C in {1,3}, S in {1,2})", C, S);
}

static int check_desc(const md2_vsl_desc* d, bool need_loss_inputs) {
    MD2_REQUIRE(d != nullptr, "null descriptor");
    MD2_REQUIRE(d->W >= 2 && d->H >= 2, "W and H must be >= 2 (reflect padding)");
    MD2_REQUIRE(d->N >= 1, "N must be >= 1");
    MD2_REQUIRE(d->C == 1 || d->C == 3, "C must be 1 or 3");
    MD2_REQUIRE(d->S >= 1 && d->S <= MAX_S, "S must be 1 or 2");
    MD2_REQUIRE(d->K && d->invK, "K / invK are null");
    for (int s = 0; s < d->S; ++s) {
        MD2_REQUIRE(d->source[s] != nullptr, "null source image");
        MD2_REQUIRE(d->rot[s] && d->trans[s], "null pose");
    }
    if (need_loss_inputs) {
        MD2_REQUIRE(d->L >= 1 && d->L <= MAX_L, "L must be in 1..8");
        MD2_REQUIRE((long long)d->L * d->N <= 65535, "L*N too large");
        MD2_REQUIRE(d->target != nullptr, "null target image");
        for (int l = 0; l < d->L; ++l) {
            MD2_REQUIRE(d->disparity[l] != nullptr, "null disparity");
            MD2_REQUIRE(d->disp_w[l] >= 2 && d->disp_h[l] >= 2 && d->disp_w[l] <= d->W && d->disp_h[l] <= d->H,
                        "bad disparity size");
        }
    }
    return 0;
}

int fill_pose_args(const md2_vsl_desc* d, PoseArgs& pa) {
    pa.S = d->S; pa.N = d->N; pa.mode = d->pose_mode;
    pa.K = d->K; pa.invK = d->invK;
    for (int s = 0; s < MAX_S; ++s) {
        pa.rot[s] = s < d->S ? d->rot[s] : nullptr;
        pa.trans[s] = s < d->S ? d->trans[s] : nullptr;
        pa.invert[s] = s < d->S ? d->invert[s] : 0;
        pa.grot[s] = s < d->S ? d->grad_rot[s] : nullptr;
        pa.gtrans[s] = s < d->S ? d->grad_trans[s] : nullptr;
    }
    return 0;
}

int prepare_pose(md2_ctx* ctx, const md2_vsl_desc* d, float** pose_ab, cudaStream_t st) {
    float* ab = (float*)ws_get(ctx, MD2_WS_POSE, sizeof(float) * 12 * d->S * d->N);
    if (!ab) return 1;
    PoseArgs pa;
    fill_pose_args(d, pa);
    pose_prep_kernel<<<cdiv(d->S * d->N, 64), 64, 0, st>>>(pa, ab);
    MD2_LAUNCH_CHECK(ctx);
    *pose_ab = ab;
    return 0;
}

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_FWDBWD = 2 };

static int run_vsl(md2_ctx* ctx, const md2_vsl_desc* d, int mode, float gloss, cudaStream_t st) {
    if (check_desc(d, true)) return 1;
    MD2_CHECK(cudaSetDevice(ctx->device));
    const int W = d->W, H = d->H, N = d->N, L = d->L, S = d->S, C = d->C;
    const long long HW = (long long)W * H;
    const bool bwd = mode != MODE_FWD;

    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.W = W; p.H = H; p.N = N; p.L = L;
    p.tgt = d->target; p.tgt_ns = d->target_image_stride;
    for (int s = 0; s < S; ++s) {
        p.src[s] = d->source[s]; p.src_ns[s] = d->source_image_stride[s];
        p.gsrc[s] = bwd ? d->grad_source[s] : nullptr;
        p.viz_warped[s] = d->viz_warped[s];
    }
    p.viz_loss = d->viz_loss;
    p.automask = d->automask;
    p.depth_a = (float)(1.0 / d->min_depth - 1.0 / d->max_depth);
    {   // the reference rounds min_disp and max_disp to T first (src/utils.jl:176-178)
        const float mind = (float)(1.0 / (double)d->max_depth), maxd = (float)(1.0 / (double)d->min_depth);
        p.depth_a = maxd - mind; p.depth_b = mind;
    }
    for (int l = 0; l < L; ++l) p.smooth_w[l] = d->smooth_weight[l];
    p.loss_scale = d->loss_scale;
    p.gloss = gloss;
    p.normalize_disp = d->normalize_disparity;

    float* pose_ab = nullptr;
    if (prepare_pose(ctx, d, &pose_ab, st)) return 1;
    p.pose_ab = pose_ab;

    // full-resolution disparities (A17: upsample_bilinear when the decoder scale is smaller)
    int n_up = 0;
    for (int l = 0; l < L; ++l) n_up += (d->disp_w[l] != W || d->disp_h[l] != H);
    float* up = nullptr; float* gup = nullptr;
    if (n_up) {
        up = (float*)ws_get(ctx, MD2_WS_DISP, sizeof(float) * n_up * N * HW);
        if (!up) return 1;
        if (bwd) {
            gup = (float*)ws_get(ctx, MD2_WS_GDISP, sizeof(float) * n_up * N * HW);
            if (!gup) return 1;
        }
    }
    for (int l = 0, k = 0; l < L; ++l) {
        if (d->disp_w[l] != W || d->disp_h[l] != H) {
            float* o = up + (long long)k * N * HW;
            upsample_kernel<<<cdiv(N * HW, 256), 256, 0, st>>>(d->disparity[l], o, d->disp_w[l], d->disp_h[l], W, H, N);
            MD2_LAUNCH_CHECK(ctx);
            p.disp[l] = o;
            p.gdisp[l] = bwd ? gup + (long long)k * N * HW : nullptr;
            ++k;
        } else {
            p.disp[l] = d->disparity[l];
            p.gdisp[l] = bwd ? d->grad_disparity[l] : nullptr;
        }
        if (bwd) MD2_REQUIRE(d->grad_disparity[l] != nullptr, "null grad_disparity");
    }

    const int tiles = cdiv(W, TILE_W) * cdiv(H, TILE_H);
    const int NP = NSTAT + 12 * S;
    float* partial = (float*)ws_get(ctx, MD2_WS_PARTIAL, sizeof(float) * (size_t)tiles * L * N * NP);
    float* sums = (float*)ws_get(ctx, MD2_WS_SUMS, sizeof(float) * (size_t)L * N * NP);
    float* stats = (float*)ws_get(ctx, MD2_WS_STATS, sizeof(float) * (size_t)L * N * NSTAT);
    if (!partial || !sums || !stats) return 1;
    p.partial = partial;

    if (mode == MODE_BWD) {
        if (d->saved) stats = d->saved;   // else: the ctx still holds the last forward's statistics
    } else if (mode == MODE_FWDBWD) {
        StatsArgs sa;
        sa.W = W; sa.H = H; sa.N = N; sa.L = L; sa.tgt = d->target; sa.tgt_ns = d->target_image_stride;
        for (int l = 0; l < MAX_L; ++l) sa.disp[l] = p.disp[l];
        const int bpi = max(1, min(64, cdiv(HW, 2048)));
        float* part2 = (float*)ws_get(ctx, MD2_WS_MISC, sizeof(float) * (size_t)bpi * L * N * NSTAT);
        if (!part2) return 1;
        dim3 g(bpi, L * N);
        if (C == 1) stats_kernel<1><<<g, 256, 0, st>>>(sa, part2);
        else stats_kernel<3><<<g, 256, 0, st>>>(sa, part2);
        MD2_LAUNCH_CHECK(ctx);
        if (launch_reduce_partials(ctx, part2, stats, NSTAT, nullptr, L * N, bpi, NSTAT, st)) return 1;
    }
    p.stats = stats;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->prof_on) {
        if (ctx->prof_used + 2 > ctx->prof_ev.size()) {
            for (int k = 0; k < 2; ++k) {
                cudaEvent_t e;
                MD2_CHECK(cudaEventCreate(&e));
                ctx->prof_ev.push_back(e);
            }
        }
        ev0 = ctx->prof_ev[ctx->prof_used]; ev1 = ctx->prof_ev[ctx->prof_used + 1];
        ctx->prof_used += 2;
        MD2_CHECK(cudaEventRecord(ev0, st));
    }
    if (bwd) { if (dispatch_fused<true>(ctx, C, S, p, st)) return 1; }
    else     { if (dispatch_fused<false>(ctx, C, S, p, st)) return 1; }
    if (ev1) MD2_CHECK(cudaEventRecord(ev1, st));
    if (launch_reduce_partials(ctx, partial, sums, NP, nullptr, L * N, tiles, NP, st)) return 1;

    FinalArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.W = W; fa.H = H; fa.N = N; fa.L = L; fa.S = S; fa.NP = NP; fa.mode = mode;
    fa.sums = sums; fa.stats = stats; fa.saved = (mode == MODE_BWD) ? nullptr : d->saved;
    if (fa.saved == stats) fa.saved = nullptr;
    fa.loss = d->loss;
    for (int l = 0; l < L; ++l) fa.smooth_w[l] = d->smooth_weight[l];
    fa.loss_scale = d->loss_scale; fa.normalize_disp = d->normalize_disparity;
    fill_pose_args(d, fa.pose);
    finalize_kernel<<<1, 128, 0, st>>>(fa);
    MD2_LAUNCH_CHECK(ctx);

    if (bwd) {
        for (int l = 0, k = 0; l < L; ++l) {
            if (d->disp_w[l] != W || d->disp_h[l] != H) {
                const long long cnt = (long long)d->disp_w[l] * d->disp_h[l] * N;
                upsample_bwd_kernel<<<cdiv(cnt, 128), 128, 0, st>>>(gup + (long long)k * N * HW, d->grad_disparity[l],
                                                                    d->disp_w[l], d->disp_h[l], W, H, N);
                MD2_LAUNCH_CHECK(ctx);
                ++k;
            }
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// A15 `warp` (called at src/simple_depth.jl:30-32, body inferred from src/training.jl:48-57):
// disparity -> S warped images, thread per pixel, and its adjoint
// ------------------------------------------------------------------------------------------
struct WarpIO {
    float* out[MAX_S];          // fwd: (N,C,H,W) contiguous
    const float* gout[MAX_S];   // bwd
};

template <int C, int S>
__global__ void __launch_bounds__(256) warp_fwd_kernel(const __grid_constant__ FusedParams p, WarpIO io) {
    const long long HW = (long long)p.W * p.H;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW * p.N) return;
    const int n = (int)(i / HW), pix = (int)(i % HW), gx = pix % p.W, gy = pix / p.W;
    const float d = p.disp[0][i];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        float val[C]; Taps tp; Proj pr; float z;
        Fused<C, S, false>::template warp_pixel<false>(p, n, s, gx, gy, d, p.pose_ab + ((long long)s * p.N + n) * 12,
                                                       val, nullptr, nullptr, tp, pr, z);
#pragma unroll
        for (int c = 0; c < C; ++c) io.out[s][((long long)n * C + c) * HW + pix] = val[c];
    }
}

template <int C, int S>
__global__ void __launch_bounds__(256) warp_bwd_kernel(const __grid_constant__ FusedParams p, WarpIO io) {
    __shared__ float scratch[(NSTAT + 12 * S) * 8];
    const long long HW = (long long)p.W * p.H;
    const int n = blockIdx.y;
    constexpr int NP = NSTAT + 12 * S;
    float v[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) v[k] = 0.f;
    for (long long pl = (long long)blockIdx.x * blockDim.x + threadIdx.x; pl < HW; pl += (long long)gridDim.x * blockDim.x) {
        const int pix = (int)pl, gx = pix % p.W, gy = pix / p.W;
        const float d = p.disp[0][(long long)n * HW + pix];
        float dbar_z = 0.f, zz = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            float val[C], dix[C], diy[C], ibar[C]; Taps tp; Proj pr; float z;
            Fused<C, S, true>::template warp_pixel<true>(p, n, s, gx, gy, d, p.pose_ab + ((long long)s * p.N + n) * 12,
                                                         val, dix, diy, tp, pr, z);
            zz = z;
            float du = 0.f, dv = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                ibar[c] = io.gout[s][((long long)n * C + c) * HW + pix];
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
                    atomicAdd(gc + tp.y0 * p.W + tp.x0, w00 * ibar[c]);
                    if (tp.x0 + 1 < p.W) atomicAdd(gc + tp.y0 * p.W + tp.x0 + 1, w01 * ibar[c]);
                    if (tp.y0 + 1 < p.H) atomicAdd(gc + (tp.y0 + 1) * p.W + tp.x0, w10 * ibar[c]);
                    if (tp.x0 + 1 < p.W && tp.y0 + 1 < p.H) atomicAdd(gc + (tp.y0 + 1) * p.W + tp.x0 + 1, w11 * ibar[c]);
                }
            }
            float cb[3];
            project_ab_bwd(pr, du, dv, cb);
            dbar_z += cb[0] * pr.ap[0] + cb[1] * pr.ap[1] + cb[2] * pr.ap[2];
            const float zp[3] = {z * (float)(gx + 1), z * (float)(gy + 1), z};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int b = 0; b < 3; ++b) v[NSTAT + 12 * s + 3 * a + b] = fmaf(cb[a], zp[b], v[NSTAT + 12 * s + 3 * a + b]);
                v[NSTAT + 12 * s + 9 + a] += cb[a];
            }
        }
        p.gdisp[0][(long long)n * HW + pix] = -p.depth_a * zz * zz * dbar_z;
    }
    block_sum<NP>(v, scratch);
    if (threadIdx.x < NP)
        p.partial[((long long)n * gridDim.x + blockIdx.x) * NP + threadIdx.x] = scratch[threadIdx.x * (blockDim.x >> 5)];
}

static int run_warp(md2_ctx* ctx, const md2_vsl_desc* d, float* const* out, const float* const* gout, cudaStream_t st) {
    if (check_desc(d, false)) return 1;
    MD2_CHECK(cudaSetDevice(ctx->device));
    const int W = d->W, H = d->H, N = d->N, S = d->S, C = d->C;
    const long long HW = (long long)W * H;
    const bool bwd = gout != nullptr;
    MD2_REQUIRE(d->disparity[0] != nullptr, "null disparity");
    MD2_REQUIRE(d->disp_w[0] == W && d->disp_h[0] == H, "warp needs a full-resolution disparity");
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.W = W; p.H = H; p.N = N; p.L = 1;
    WarpIO io;
    memset(&io, 0, sizeof(io));
    for (int s = 0; s < S; ++s) {
        p.src[s] = d->source[s]; p.src_ns[s] = d->source_image_stride[s];
        p.gsrc[s] = bwd ? d->grad_source[s] : nullptr;
        if (bwd) { MD2_REQUIRE(gout[s] != nullptr, "null upstream gradient"); io.gout[s] = gout[s]; }
        else { MD2_REQUIRE(out && out[s], "null output"); io.out[s] = out[s]; }
    }
    const float mind = (float)(1.0 / (double)d->max_depth), maxd = (float)(1.0 / (double)d->min_depth);
    p.depth_a = maxd - mind; p.depth_b = mind;
    p.disp[0] = d->disparity[0];
    float* pose_ab = nullptr;
    if (prepare_pose(ctx, d, &pose_ab, st)) return 1;
    p.pose_ab = pose_ab;
    if (!bwd) {
        const int g = cdiv(HW * N, 256);
        if (C == 1 && S == 1) warp_fwd_kernel<1, 1><<<g, 256, 0, st>>>(p, io);
        else if (C == 1 && S == 2) warp_fwd_kernel<1, 2><<<g, 256, 0, st>>>(p, io);
        else if (C == 3 && S == 1) warp_fwd_kernel<3, 1><<<g, 256, 0, st>>>(p, io);
        else warp_fwd_kernel<3, 2><<<g, 256, 0, st>>>(p, io);
        MD2_LAUNCH_CHECK(ctx);
        return 0;
    }
    MD2_REQUIRE(d->grad_disparity[0] != nullptr, "null grad_disparity");
    p.gdisp[0] = d->grad_disparity[0];
    const int NP = NSTAT + 12 * S;
    const int bpi = max(1, min(128, cdiv(HW, 1024)));
    float* partial = (float*)ws_get(ctx, MD2_WS_PARTIAL, sizeof(float) * (size_t)bpi * N * NP);
    float* sums = (float*)ws_get(ctx, MD2_WS_SUMS, sizeof(float) * (size_t)N * NP);
    if (!partial || !sums) return 1;
    p.partial = partial;
    dim3 g(bpi, N);
    if (C == 1 && S == 1) warp_bwd_kernel<1, 1><<<g, 256, 0, st>>>(p, io);
    else if (C == 1 && S == 2) warp_bwd_kernel<1, 2><<<g, 256, 0, st>>>(p, io);
    else if (C == 3 && S == 1) warp_bwd_kernel<3, 1><<<g, 256, 0, st>>>(p, io);
    else warp_bwd_kernel<3, 2><<<g, 256, 0, st>>>(p, io);
    MD2_LAUNCH_CHECK(ctx);
    if (launch_reduce_partials(ctx, partial, sums, NP, nullptr, N, bpi, NP, st)) return 1;
    FinalArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.W = W; fa.H = H; fa.N = N; fa.L = 1; fa.S = S; fa.NP = NP; fa.mode = MODE_BWD;
    fa.sums = sums;
    fill_pose_args(d, fa.pose);
    finalize_kernel<<<1, 128, 0, st>>>(fa);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace md2

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char* md2_version(void) { return "md2_b200 0.1.0 (sm_100a)"; }
const char* md2_last_error(void) { return md2::last_error_ref().c_str(); }

int md2_create(int device, md2_ctx** out) {
    if (!out) return md2::set_error("md2_create: null out pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return md2::set_error("md2_create: no CUDA device (%s); this library has no CPU fallback",
                              cudaGetErrorString(e));
    if (device < 0 || device >= count) return md2::set_error("md2_create: bad device %d", device);
    MD2_CHECK(cudaSetDevice(device));
    md2_ctx* c = new md2_ctx();
    c->device = device;
    c->launches = 0;
    *out = c;
    return 0;
}

int md2_destroy(md2_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < MD2_WS_COUNT; ++i)
        if (ctx->ws[i].ptr) cudaFree(ctx->ws[i].ptr);
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    delete ctx;
    return 0;
}

int64_t md2_launch_count(const md2_ctx* ctx) { return ctx ? ctx->launches : 0; }

int md2_profile_enable(md2_ctx* ctx, int32_t on) {
    MD2_REQUIRE(ctx != nullptr, "null ctx");
    ctx->prof_on = on;
    ctx->prof_used = 0;
    return 0;
}
int md2_profile_read(md2_ctx* ctx, float* total_ms, int64_t* launches) {
    MD2_REQUIRE(ctx != nullptr && total_ms && launches, "bad arguments");
    float tot = 0.f;
    for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
        MD2_CHECK(cudaEventSynchronize(ctx->prof_ev[i + 1]));
        float ms = 0.f;
        MD2_CHECK(cudaEventElapsedTime(&ms, ctx->prof_ev[i], ctx->prof_ev[i + 1]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = (int64_t)(ctx->prof_used / 2);
    ctx->prof_used = 0;
    return 0;
}

int md2_view_synthesis_loss_fwd(md2_ctx* ctx, const md2_vsl_desc* d, md2_stream st) {
    MD2_REQUIRE(ctx != nullptr, "null ctx");
    return md2::run_vsl(ctx, d, md2::MODE_FWD, 0.f, (cudaStream_t)st);
}
int md2_view_synthesis_loss_bwd(md2_ctx* ctx, const md2_vsl_desc* d, float upstream, md2_stream st) {
    MD2_REQUIRE(ctx != nullptr, "null ctx");
    return md2::run_vsl(ctx, d, md2::MODE_BWD, upstream, (cudaStream_t)st);
}
int md2_view_synthesis_loss_fwdbwd(md2_ctx* ctx, const md2_vsl_desc* d, float seed, md2_stream st) {
    MD2_REQUIRE(ctx != nullptr, "null ctx");
    return md2::run_vsl(ctx, d, md2::MODE_FWDBWD, seed, (cudaStream_t)st);
}

int md2_warp_fwd(md2_ctx* ctx, const md2_vsl_desc* d, float* const* out, md2_stream st) {
    MD2_REQUIRE(ctx != nullptr, "null ctx");
    return md2::run_warp(ctx, d, out, nullptr, (cudaStream_t)st);
}
int md2_warp_bwd(md2_ctx* ctx, const md2_vsl_desc* d, const float* const* gout, md2_stream st) {
    MD2_REQUIRE(ctx != nullptr, "null ctx");
    MD2_REQUIRE(gout != nullptr, "null upstream gradients");
    return md2::run_warp(ctx, d, nullptr, gout, (cudaStream_t)st);
}

int md2_upsample_bilinear_fwd(md2_ctx* ctx, const float* in, float* out, int32_t w, int32_t h, int32_t W, int32_t H,
                              int32_t CN, md2_stream st) {
    MD2_REQUIRE(ctx != nullptr, "null ctx");
    MD2_CHECK(cudaSetDevice(ctx->device));
    MD2_REQUIRE(in && out && w > 0 && h > 0 && W > 0 && H > 0 && CN > 0, "bad arguments");
    md2::upsample_kernel<<<md2::cdiv((long long)W * H * CN, 256), 256, 0, (cudaStream_t)st>>>(in, out, w, h, W, H, CN);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_upsample_bilinear_bwd(md2_ctx* ctx, const float* gout, float* gin, int32_t w, int32_t h, int32_t W, int32_t H,
                              int32_t CN, md2_stream st) {
    MD2_REQUIRE(ctx != nullptr, "null ctx");
    MD2_CHECK(cudaSetDevice(ctx->device));
    MD2_REQUIRE(gout && gin && w > 0 && h > 0 && W > 0 && H > 0 && CN > 0, "bad arguments");
    md2::upsample_bwd_kernel<<<md2::cdiv((long long)w * h * CN, 128), 128, 0, (cudaStream_t)st>>>(gout, gin, w, h, W, H, CN);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // extern "C"
