// Stand-alone operators of the path (one reference function each, forward + backward) so that
// every entry point of src/utils.jl and src/training.jl:1-19 has a drop-in.  These are the simple
// thread-per-element versions; the speed path is the fused kernel in md2_fused.cu.
#include <string.h>

#include "md2_common.cuh"
#include "md2_fused.cuh"

namespace md2 {

int launch_reduce_partials(md2_ctx* ctx, const float* partial, float* out0, int n0, float* out1,
                           int groups, int bpg, int NP, cudaStream_t st);

// ------------------------------------------------------------------------------------------
// A1 disparity_to_depth (src/utils.jl:175-179)
// ------------------------------------------------------------------------------------------
__global__ void d2d_fwd_kernel(const float* __restrict__ d, float* __restrict__ z, long long n, float a, float b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = 1.0f / fmaf(d[i], a, b);
}
__global__ void d2d_bwd_kernel(const float* __restrict__ d, const float* __restrict__ gz, float* __restrict__ gd,
                               long long n, float a, float b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float z = 1.0f / fmaf(d[i], a, b);
        gd[i] = -a * z * z * gz[i];
    }
}

// ------------------------------------------------------------------------------------------
// A2 Backproject (src/utils.jl:41-65)
// ------------------------------------------------------------------------------------------
__global__ void backproject_fwd_kernel(const float* __restrict__ depth, const float* __restrict__ invK,
                                       float* __restrict__ pts, int W, int H, int N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long P = (long long)W * H;
    if (i >= P * N) return;
    const long long p = i % P;
    const float w = (float)(p % W + 1), h = (float)(p / W + 1);
    const float z = depth[i];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float ray = fmaf(invK[r], w, fmaf(invK[3 + r], h, invK[6 + r]));
        pts[3 * i + r] = z * ray;
    }
}
__global__ void backproject_bwd_kernel(const float* __restrict__ gpts, const float* __restrict__ invK,
                                       float* __restrict__ gdepth, int W, int H, int N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long P = (long long)W * H;
    if (i >= P * N) return;
    const long long p = i % P;
    const float w = (float)(p % W + 1), h = (float)(p / W + 1);
    float g = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float ray = fmaf(invK[r], w, fmaf(invK[3 + r], h, invK[6 + r]));
        g = fmaf(ray, gpts[3 * i + r], g);
    }
    gdepth[i] = g;
}

// ------------------------------------------------------------------------------------------
// A3 Project + normalize (src/utils.jl:67-99).  K, R column-major: m[3*col+row]
// ------------------------------------------------------------------------------------------
struct ProjPoint { float Y[3], c[3], q, u, v; };

__device__ __forceinline__ void project_point(const float* X, const float* K, const float* R, const float* t,
                                              ProjPoint& o) {
#pragma unroll
    for (int r = 0; r < 3; ++r) o.Y[r] = fmaf(R[r], X[0], fmaf(R[3 + r], X[1], fmaf(R[6 + r], X[2], t[r])));
#pragma unroll
    for (int r = 0; r < 3; ++r) o.c[r] = fmaf(K[r], o.Y[0], fmaf(K[3 + r], o.Y[1], K[6 + r] * o.Y[2]));
    o.q = 1.0f / (o.c[2] + PROJ_EPS);
    o.u = o.c[0] * o.q;
    o.v = o.c[1] * o.q;
}

__global__ void project_fwd_kernel(const float* __restrict__ pts, const float* __restrict__ K,
                                   const float* __restrict__ R, const float* __restrict__ t,
                                   float* __restrict__ uv, int W, int H, int N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long P = (long long)W * H;
    if (i >= P * N) return;
    const int n = (int)(i / P);
    const float X[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    ProjPoint o;
    project_point(X, K, R + 9 * n, t + 3 * n, o);
    uv[2 * i] = ((o.u - 1.0f) / (float)(W - 1) - 0.5f) * 2.0f;
    uv[2 * i + 1] = ((o.v - 1.0f) / (float)(H - 1) - 0.5f) * 2.0f;
}

__global__ void __launch_bounds__(256) project_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ K,
                                                          const float* __restrict__ R, const float* __restrict__ t,
                                                          const float* __restrict__ guv, float* __restrict__ gpts,
                                                          float* __restrict__ partial, int W, int H, int N) {
    __shared__ float scratch[12 * 8];
    const int n = blockIdx.y;
    const long long P = (long long)W * H;
    const float* Rn = R + 9 * n;
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const long long i = (long long)n * P + p;
        const float X[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
        ProjPoint o;
        project_point(X, K, Rn, t + 3 * n, o);
        const float du = guv[2 * i] * 2.0f / (float)(W - 1), dv = guv[2 * i + 1] * 2.0f / (float)(H - 1);
        const float cb[3] = {du * o.q, dv * o.q, -(du * o.u + dv * o.v) * o.q};
        float Yb[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) Yb[r] = K[3 * r] * cb[0] + K[3 * r + 1] * cb[1] + K[3 * r + 2] * cb[2];  // K^T cb
#pragma unroll
        for (int r = 0; r < 3; ++r) gpts[3 * i + r] = Rn[3 * r] * Yb[0] + Rn[3 * r + 1] * Yb[1] + Rn[3 * r + 2] * Yb[2];  // R^T Yb
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) v[3 * c + r] = fmaf(Yb[r], X[c], v[3 * c + r]);   // column-major gR
#pragma unroll
        for (int r = 0; r < 3; ++r) v[9 + r] += Yb[r];
    }
    block_sum<12>(v, scratch);
    if (threadIdx.x < 12) {
        float* o = partial + ((long long)n * gridDim.x + blockIdx.x) * 12;
        o[threadIdx.x] = scratch[threadIdx.x * (blockDim.x >> 5)];
    }
}

// ------------------------------------------------------------------------------------------
// A4-A6 so3_exp_map / hat / composeT (src/utils.jl:101-141, 181-188).  Column-major I/O.
// ------------------------------------------------------------------------------------------
__global__ void so3_fwd_kernel(const float* __restrict__ rvec, float* __restrict__ R, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double r[3] = {rvec[3 * n], rvec[3 * n + 1], rvec[3 * n + 2]}, M[9];
    so3_exp(r, M);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[9 * n + 3 * j + i] = (float)M[3 * i + j];
}
__global__ void so3_bwd_kernel(const float* __restrict__ rvec, const float* __restrict__ gR, float* __restrict__ grvec, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double r[3] = {rvec[3 * n], rvec[3 * n + 1], rvec[3 * n + 2]}, Rb[9], rb[3];
    load_cm3(gR + 9 * n, Rb);
    so3_exp_bwd(r, Rb, rb);
    for (int k = 0; k < 3; ++k) grvec[3 * n + k] = (float)rb[k];
}
__global__ void hat_fwd_kernel(const float* __restrict__ rvec, float* __restrict__ S, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float a = rvec[3 * n], b = rvec[3 * n + 1], c = rvec[3 * n + 2];
    float* o = S + 9 * n;   // column-major: o[3*col+row]
    o[0] = 0.f; o[1] = c;   o[2] = -b;
    o[3] = -c;  o[4] = 0.f; o[5] = a;
    o[6] = b;   o[7] = -a;  o[8] = 0.f;
}
__global__ void hat_bwd_kernel(const float* __restrict__ gS, float* __restrict__ grvec, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* d = gS + 9 * n;   // d[3*col+row]
    grvec[3 * n + 0] = d[3 * 1 + 2] - d[3 * 2 + 1];   // d[3,2] - d[2,3]
    grvec[3 * n + 1] = -d[3 * 0 + 2] + d[3 * 2 + 0];  // -d[3,1] + d[1,3]
    grvec[3 * n + 2] = d[3 * 0 + 1] - d[3 * 1 + 0];   // d[2,1] - d[1,2]
}
__global__ void compose_fwd_kernel(const float* __restrict__ rvec, const float* __restrict__ tvec, int invert,
                                   float* __restrict__ R, float* __restrict__ t, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double r[3], tv[3], M[9], tu[3];
    for (int k = 0; k < 3; ++k) { r[k] = rvec[3 * n + k]; tv[k] = tvec[3 * n + k]; }
    compose_T(r, tv, invert, M, tu);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[9 * n + 3 * j + i] = (float)M[3 * i + j];
    for (int k = 0; k < 3; ++k) t[3 * n + k] = (float)tu[k];
}
__global__ void compose_bwd_kernel(const float* __restrict__ rvec, const float* __restrict__ tvec, int invert,
                                   const float* __restrict__ gR, const float* __restrict__ gt,
                                   float* __restrict__ grvec, float* __restrict__ gtvec, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double r[3], tv[3], Rub[9], tub[3], rb[3], tb[3];
    for (int k = 0; k < 3; ++k) { r[k] = rvec[3 * n + k]; tv[k] = tvec[3 * n + k]; tub[k] = gt ? gt[3 * n + k] : 0.0; }
    if (gR) load_cm3(gR + 9 * n, Rub);
    else for (int k = 0; k < 9; ++k) Rub[k] = 0.0;
    compose_T_bwd(r, tv, invert, Rub, tub, rb, tb);
    for (int k = 0; k < 3; ++k) { grvec[3 * n + k] = (float)rb[k]; gtvec[3 * n + k] = (float)tb[k]; }
}

// ------------------------------------------------------------------------------------------
// A16 NNlib.grid_sample bilinear / align-corners, padding :zeros or :border, and its adjoint
// ------------------------------------------------------------------------------------------
struct GsTaps { int x0, y0; float ix, iy, mx, my; };

__device__ __forceinline__ GsTaps gs_taps(float gx, float gy, int W, int H, int border) {
    GsTaps t;
    float ix = (gx + 1.0f) * 0.5f * (float)(W - 1);   // 0-based un-normalised
    float iy = (gy + 1.0f) * 0.5f * (float)(H - 1);
    t.mx = 0.5f * (float)(W - 1);
    t.my = 0.5f * (float)(H - 1);
    if (border) {
        if (!(ix > 0.f)) { ix = 0.f; t.mx = 0.f; } else if (ix >= (float)(W - 1)) { ix = (float)(W - 1); t.mx = 0.f; }
        if (!(iy > 0.f)) { iy = 0.f; t.my = 0.f; } else if (iy >= (float)(H - 1)) { iy = (float)(H - 1); t.my = 0.f; }
    } else {
        // keep the int conversion defined for wild coordinates; such taps are out of range anyway
        ix = fminf(fmaxf(ix, -2.0f), (float)W + 1.0f);
        iy = fminf(fmaxf(iy, -2.0f), (float)H + 1.0f);
    }
    t.ix = ix; t.iy = iy;
    t.x0 = (int)floorf(ix); t.y0 = (int)floorf(iy);
    return t;
}

__global__ void grid_sample_fwd_kernel(const float* __restrict__ in, const float* __restrict__ grid,
                                       float* __restrict__ out, int W, int H, int C, int N, int Wo, int Ho, int border) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long Po = (long long)Wo * Ho;
    if (i >= Po * N) return;
    const int n = (int)(i / Po);
    const long long po = i % Po;
    const GsTaps t = gs_taps(grid[2 * i], grid[2 * i + 1], W, H, border);
    const float fx = t.ix - (float)t.x0, fy = t.iy - (float)t.y0;
    const bool xa = t.x0 >= 0 && t.x0 < W, xb = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    const bool ya = t.y0 >= 0 && t.y0 < H, yb = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    for (int c = 0; c < C; ++c) {
        const float* b = in + ((long long)n * C + c) * W * H;
        const float v00 = (xa && ya) ? b[t.y0 * W + t.x0] : 0.f;
        const float v01 = (xb && ya) ? b[t.y0 * W + t.x0 + 1] : 0.f;
        const float v10 = (xa && yb) ? b[(t.y0 + 1) * W + t.x0] : 0.f;
        const float v11 = (xb && yb) ? b[(t.y0 + 1) * W + t.x0 + 1] : 0.f;
        out[((long long)n * C + c) * Po + po] =
            v00 * (1.f - fx) * (1.f - fy) + v01 * fx * (1.f - fy) + v10 * (1.f - fx) * fy + v11 * fx * fy;
    }
}

__global__ void grid_sample_bwd_kernel(const float* __restrict__ in, const float* __restrict__ grid,
                                       const float* __restrict__ gout, float* __restrict__ gin,
                                       float* __restrict__ ggrid, int W, int H, int C, int N, int Wo, int Ho, int border) {
    // (no early exit: every lane of a warp takes part in the shuffles of the merged scatter)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long Po = (long long)Wo * Ho;
    const bool active = i < Po * N;
    const long long ii = active ? i : 0;
    const int lane = threadIdx.x & 31;
    const int n = (int)(ii / Po);
    const long long po = ii % Po;
    const GsTaps t = gs_taps(grid[2 * ii], grid[2 * ii + 1], W, H, border);
    const float fx = t.ix - (float)t.x0, fy = t.iy - (float)t.y0;
    const bool xa = t.x0 >= 0 && t.x0 < W, xb = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    const bool ya = t.y0 >= 0 && t.y0 < H, yb = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    float gix = 0.f, giy = 0.f;
    for (int c = 0; c < C; ++c) {
        const long long pl = ((long long)n * C + c) * W * H;
        const float* b = in + pl;
        const float g = active ? gout[((long long)n * C + c) * Po + po] : 0.f;
        const float v00 = (xa && ya) ? b[t.y0 * W + t.x0] : 0.f;
        const float v01 = (xb && ya) ? b[t.y0 * W + t.x0 + 1] : 0.f;
        const float v10 = (xa && yb) ? b[(t.y0 + 1) * W + t.x0] : 0.f;
        const float v11 = (xb && yb) ? b[(t.y0 + 1) * W + t.x0 + 1] : 0.f;
        gix += g * ((v01 - v00) * (1.f - fy) + (v11 - v10) * fy);
        giy += g * ((v10 - v00) * (1.f - fx) + (v11 - v01) * fx);
        if (gin) {   // (uniform: a kernel argument)
            float* r0 = gin + pl + (long long)t.y0 * W + t.x0;
            float* r1 = r0 + W;
            const bool top = active && ya, bot = active && yb;
            red_pair_merged((top && xa) ? r0 : nullptr, (top && xb) ? r0 + 1 : nullptr, g * (1.f - fx) * (1.f - fy), g * fx * (1.f - fy), lane);
            red_pair_merged((bot && xa) ? r1 : nullptr, (bot && xb) ? r1 + 1 : nullptr, g * (1.f - fx) * fy, g * fx * fy, lane);
        }
    }
    if (ggrid && active) {
        ggrid[2 * i] = t.mx * gix;
        ggrid[2 * i + 1] = t.my * giy;
    }
}

// ------------------------------------------------------------------------------------------
// A7 SSIM (src/utils.jl:13-39) and A10-A13 photometric / min / mask (src/training.jl:1-19).
// Gather formulation: deterministic, no atomics.
// ------------------------------------------------------------------------------------------
struct WinXY {           // SSIM window statistics with the coefficients for both arguments
    float s, pass;
    float ax, bx, g;     // dS/dx_j = ax + bx x_j + g y_j
    float ay, by;        // dS/dy_j = ay + by y_j + g x_j
};

// statistics of one window from its nine samples of x and y (index 4 = the centre pixel)
__device__ __forceinline__ WinXY window_core(const float (&xv)[9], const float (&yv)[9]) {
    const float xc = xv[4], yc = yv[4];
    float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float a = xv[k] - xc, b = yv[k] - yc;
        sx += a; sy += b;
        sxx = fmaf(a, a, sxx); syy = fmaf(b, b, syy); sxy = fmaf(a, b, sxy);
    }
    const float r9 = 1.0f / 9.0f;
    const float dx = sx * r9, dy = sy * r9;
    const float mux = xc + dx, muy = yc + dy;
    const float vx = fmaf(-dx, dx, sxx * r9), vy = fmaf(-dy, dy, syy * r9), vxy = fmaf(-dx, dy, sxy * r9);
    const float A = 2.0f * mux * muy + SSIM_C1, B = 2.0f * vxy + SSIM_C2;
    const float Cc = (mux * mux + muy * muy) + SSIM_C1, D = (vx + vy) + SSIM_C2;
    const float S = (A * B) / (Cc * D);   // identical inputs: numerator == denominator bit for bit
    const float rC = 1.0f / Cc, rD = 1.0f / D, inv = rC * rD;
    const float raw = (1.0f - S) * 0.5f;
    WinXY o;
    o.s = fminf(fmaxf(raw, 0.f), 1.f);
    o.pass = (raw >= 0.f && raw <= 1.f) ? 1.f : 0.f;
    const float k = 2.0f / 9.0f;
    o.bx = -k * S * rD; o.by = o.bx;
    o.g = k * A * inv;
    o.ax = k * (muy * (B - A) * inv - S * mux * rC + S * mux * rD);
    o.ay = k * (mux * (B - A) * inv - S * muy * rC + S * muy * rD);
    return o;
}

// ---- shared-memory halo tiles (the SSIM window is a 3x3 stencil; its backward reaches two pixels out) ----------
// A block owns a TW x TH tile of one image plane.  load_tile stages the tile plus a halo of HALO pixels of a plane into
// shared memory, coalesced row by row; coordinates outside the image are folded back by the reflect-pad(1) of
// src/utils.jl:26-27 (only -1 and n are ever used by an in-image window; anything further is clamped and never read).
constexpr int TILE_W = 32, TILE_H = 8, TILE_THREADS = TILE_W * TILE_H;

__device__ __forceinline__ int fold_coord(int g, int n) {
    g = reflect1(g, n);
    return g < 0 ? 0 : (g > n - 1 ? n - 1 : g);
}
// (32 x 8 threads: a thread keeps its column, so the reflect / clamp of the column index is formed once per call and the
// loop over the rows is one address, one load and one store per element; the 2 * HALO extra columns go to the first lanes)
template <int HALO, int TH = TILE_H>
__device__ __forceinline__ void load_tile(float* __restrict__ sm, const float* __restrict__ plane, int x0, int y0, int W, int H) {
    constexpr int SW = TILE_W + 2 * HALO, SH = TH + 2 * HALO;
    const int tx = threadIdx.x % TILE_W, tr = threadIdx.x / TILE_W;
    const int ca = fold_coord(x0 - HALO + tx, W);
    const int cb = fold_coord(x0 - HALO + TILE_W + tx, W);     // (used by lanes 0 .. 2 * HALO - 1)
#pragma unroll
    for (int ty = tr; ty < SH; ty += TILE_THREADS / TILE_W) {
        const float* row = plane + (long long)fold_coord(y0 - HALO + ty, H) * W;
        sm[ty * SW + tx] = row[ca];
        if (tx < 2 * HALO) sm[ty * SW + TILE_W + tx] = row[cb];
    }
}
// the nine samples of the window centred on tile-local (cx, cy) of a staged tile with row pitch SW
template <int SW>
__device__ __forceinline__ void window_samples(const float* __restrict__ sm, int cx, int cy, float (&v)[9]) {
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) v[(dy + 1) * 3 + dx + 1] = sm[(cy + dy) * SW + cx + dx];
}

__device__ __forceinline__ float fold_w(int g, int d, int n) {   // adjoint of reflect-pad(1)
    return 1.0f + ((g == 1 && d == -1) ? 1.f : 0.f) + ((g == n - 2 && d == 1) ? 1.f : 0.f);
}

__global__ void __launch_bounds__(TILE_THREADS) ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                                                                int W, int H) {
    constexpr int SW = TILE_W + 2, SH = TILE_H + 2;
    __shared__ float xs[SW * SH], ys[SW * SH];
    const long long HW = (long long)W * H, pl = blockIdx.z;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    load_tile<1>(xs, x + pl * HW, x0, y0, W, H);
    load_tile<1>(ys, y + pl * HW, x0, y0, W, H);
    __syncthreads();
    const int tx = threadIdx.x % TILE_W, ty = threadIdx.x / TILE_W;
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= W || gy >= H) return;
    float xv[9], yv[9];
    window_samples<SW>(xs, tx + 1, ty + 1, xv);
    window_samples<SW>(ys, tx + 1, ty + 1, yv);
    out[pl * HW + (long long)gy * W + gx] = window_core(xv, yv).s;
}

// Backward in two phases per tile: (1) the coefficients of every window that touches the tile (tile + halo 1), each
// computed ONCE from the staged pixels (tile + halo 2) and scaled by its upstream cotangent; (2) every pixel gathers its
// nine windows (gather form: deterministic, no atomics).  The naive form recomputes each window nine times from global memory.
__global__ void __launch_bounds__(TILE_THREADS) ssim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gout,
                                                                float* __restrict__ gx, float* __restrict__ gy, int W, int H) {
    constexpr int SW = TILE_W + 4, SH = TILE_H + 4, CW = TILE_W + 2, CH = TILE_H + 2;
    __shared__ float xs[SW * SH], ys[SW * SH];
    __shared__ float cf[5][CW * CH];                 // k * (ax, bx, g, ay, by) of the windows around the tile
    const long long HW = (long long)W * H, pl = blockIdx.z;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    load_tile<2>(xs, x + pl * HW, x0, y0, W, H);
    load_tile<2>(ys, y + pl * HW, x0, y0, W, H);
    __syncthreads();
    for (int k = threadIdx.x; k < CW * CH; k += TILE_THREADS) {
        const int wy = k / CW, wx = k - wy * CW;
        const int qx = x0 - 1 + wx, qy = y0 - 1 + wy;     // window centre in the image
        float c[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (qx >= 0 && qx < W && qy >= 0 && qy < H) {
            float xv[9], yv[9];
            window_samples<SW>(xs, wx + 1, wy + 1, xv);
            window_samples<SW>(ys, wx + 1, wy + 1, yv);
            const WinXY w = window_core(xv, yv);
            const float kk = gout[pl * HW + (long long)qy * W + qx] * (-0.5f) * w.pass;
            c[0] = kk * w.ax; c[1] = kk * w.bx; c[2] = kk * w.g; c[3] = kk * w.ay; c[4] = kk * w.by;
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) cf[j][k] = c[j];
    }
    __syncthreads();
    const int tx = threadIdx.x % TILE_W, ty = threadIdx.x / TILE_W;
    const int jx = x0 + tx, jy = y0 + ty;
    if (jx >= W || jy >= H) return;
    const float xj = xs[(ty + 2) * SW + tx + 2], yj = ys[(ty + 2) * SW + tx + 2];
    float ax = 0.f, ay = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int k = (ty + 1 + dy) * CW + tx + 1 + dx;      // (windows outside the image hold zeros)
            const float f = fold_w(jx, dx, W) * fold_w(jy, dy, H);
            ax = fmaf(f, cf[0][k] + cf[1][k] * xj + cf[2][k] * yj, ax);
            ay = fmaf(f, cf[3][k] + cf[4][k] * yj + cf[2][k] * xj, ay);
        }
    const long long o = pl * HW + (long long)jy * W + jx;
    if (gx) gx[o] = ax;
    if (gy) gy[o] = ay;
}

struct PmArgs {
    int S, W, H, C, N;
    const float* pred[MAX_S]; long long pred_ns[MAX_S];
    const float* target; long long target_ns;
    const float* mask;
    float alpha;
    float* gpred[MAX_S];
};

// out = min([mask,] pe_0, ..., pe_{S-1}), pe_s = alpha mean_c SSIM(pred_s, target) + (1 - alpha) mean_c |target - pred_s|.
// One block per 32 x 32 tile and image: 32 columns x 8 threads, every thread marches down FWD_RPT = 4 rows of its column
// with the horizontal 3-sums of a row formed once (six shared-memory reads) and rolled over three rows in registers --
// a third of the reads and less than half of the arithmetic of gathering nine samples per window and output.  The
// target tile of a channel is staged once and shared by the S sources.  (This is the automask pre-pass of every
// training step with automasking, md2_vsl_desc.compute_automask.)
constexpr int FWD_RPT = 4, FWD_TH = 8 * FWD_RPT;

// SSIM of one window from its nine-sample sums of a = x - x0, b = y - y0 (any constants x0, y0 near the window: the
// centring only keeps the squares small), src/utils.jl:25-38
__device__ __forceinline__ float ssim_from_sums(float x0, float y0, float sa, float sb, float saa, float sbb, float sab) {
    const float r9 = 1.0f / 9.0f;
    const float da = sa * r9, db = sb * r9;
    const float mux = x0 + da, muy = y0 + db;
    const float vx = fmaf(-da, da, saa * r9), vy = fmaf(-db, db, sbb * r9), vxy = fmaf(-da, db, sab * r9);
    const float A = 2.0f * mux * muy + SSIM_C1, B = 2.0f * vxy + SSIM_C2;
    const float Cc = (mux * mux + muy * muy) + SSIM_C1, D = (vx + vy) + SSIM_C2;
    const float S = (A * B) / (Cc * D);   // identical inputs: numerator == denominator bit for bit
    return fminf(fmaxf((1.0f - S) * 0.5f, 0.f), 1.f);
}

template <int C>
__global__ void __launch_bounds__(TILE_THREADS) photomin_fwd_kernel(PmArgs a, float* __restrict__ out, int* __restrict__ argmin) {
    constexpr int SW = TILE_W + 2, SH = FWD_TH + 2;
    __shared__ float ys[SW * SH], xs[SW * SH];
    const long long HW = (long long)a.W * a.H;
    const int n = blockIdx.z, x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * FWD_TH;
    const int tx = threadIdx.x % TILE_W, tr = threadIdx.x / TILE_W;          // column, group of FWD_RPT rows
    const int gx = x0 + tx, gy0 = y0 + tr * FWD_RPT;
    const float* y = a.target + n * a.target_ns;
    float ss[MAX_S][FWD_RPT], l1[MAX_S][FWD_RPT];
#pragma unroll
    for (int s = 0; s < MAX_S; ++s)
#pragma unroll
        for (int r = 0; r < FWD_RPT; ++r) { ss[s][r] = 0.f; l1[s][r] = 0.f; }
    auto stage = [&](float* sm, const float* plane) { load_tile<1, FWD_TH>(sm, plane, x0, y0, a.W, a.H); };   // tile + halo 1
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
        __syncthreads();
        stage(ys, y + c * HW);
#pragma unroll
        for (int s = 0; s < MAX_S; ++s) {
            if (s >= a.S) break;
            __syncthreads();
            stage(xs, a.pred[s] + n * a.pred_ns[s] + c * HW);
            __syncthreads();
            // centring constants of this thread's windows: the first centre sample
            const int cbase = (tr * FWD_RPT + 1) * SW + tx + 1;
            const float xr = xs[cbase], yr = ys[cbase];
            float ha[3], hb[3], haa[3], hbb[3], hab[3], yc[3], xc[3];
#pragma unroll
            for (int j = 0; j < FWD_RPT + 2; ++j) {              // tile-local rows tr*RPT + j (image rows gy0 - 1 + j)
                const int o = (tr * FWD_RPT + j) * SW + tx;
                const float al = xs[o] - xr, ac = xs[o + 1] - xr, ar = xs[o + 2] - xr;
                const float bl = ys[o] - yr, bc = ys[o + 1] - yr, br = ys[o + 2] - yr;
                const int k = j % 3;
                ha[k] = al + ac + ar; hb[k] = bl + bc + br;
                haa[k] = fmaf(ar, ar, fmaf(ac, ac, al * al)); hbb[k] = fmaf(br, br, fmaf(bc, bc, bl * bl));
                hab[k] = fmaf(ar, br, fmaf(ac, bc, al * bl));
                xc[k] = xs[o + 1]; yc[k] = ys[o + 1];
                if (j >= 2) {                                    // window centred on row j - 1 is complete
                    const int m = (j - 1) % 3;
                    ss[s][j - 2] += ssim_from_sums(xr, yr, ha[0] + ha[1] + ha[2], hb[0] + hb[1] + hb[2], haa[0] + haa[1] + haa[2],
                                                   hbb[0] + hbb[1] + hbb[2], hab[0] + hab[1] + hab[2]);
                    l1[s][j - 2] += fabsf(yc[m] - xc[m]);
                }
            }
        }
    }
    if (gx >= a.W) return;
#pragma unroll
    for (int r = 0; r < FWD_RPT; ++r) {
        const int gy = gy0 + r;
        if (gy >= a.H) break;
        const long long i = (long long)n * HW + (long long)gy * a.W + gx;
        float best = 0.f; int bi = -1;
        if (a.mask) best = a.mask[i];
#pragma unroll
        for (int s = 0; s < MAX_S; ++s) {
            if (s >= a.S) break;
            const float pe = a.alpha * (ss[s][r] * (1.0f / C)) + (1.0f - a.alpha) * (l1[s][r] * (1.0f / C));
            if ((s == 0 && !a.mask) || pe < best) { best = pe; bi = s; }
        }
        out[i] = best;
        if (argmin) argmin[i] = bi;
    }
}

// ---- the same operator with the tiles staged by the copy engine (TMA bulk copies, sm_90+) ------------------------------
// Interior tiles (x0 >= 4, x0 + 36 <= W, rows 16-byte aligned): every tile row is ONE cp.async.bulk of 40 floats (the 34
// needed + the 16-byte alignment margin) issued by one thread per row, completion counted on an mbarrier; no per-element
// address arithmetic, no register round trip, and the (channel, source) passes are double-buffered: the copies of pass
// k + 1 are in flight while pass k computes (one __syncthreads per pass instead of three).  Tiles on the left / right image
// border (reflected columns) and unaligned shapes stage through the threads into the same layout.
#ifndef MD2_PM_BULK
#define MD2_PM_BULK 1
#endif
#ifndef MD2_PM_MINB
#define MD2_PM_MINB 4      // resident blocks per SM the register allocation aims at (4: <= 64 registers)
#endif
constexpr int PM_PITCH = 40, PM_COL0 = 3;              // smem pitch (floats); tile-local column of image column x0 - 1
constexpr int PM_ROWS = FWD_TH + 2;
constexpr unsigned PM_ROW_BYTES = PM_PITCH * 4, PM_TILE_BYTES = PM_ROWS * PM_ROW_BYTES;

__device__ __forceinline__ unsigned pm_smem(const void* q) { return (unsigned)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void pm_bar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PM_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PM_DONE_%=;\n"
        "bra PM_WAIT_%=;\n"
        "PM_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// All threads call.  Thread 0 arms the barrier with the byte count; lane 0 of every warp issues the copies of tile rows
// warp, warp + 8, ... (a bulk copy takes warp-uniform operands: issued from the lanes of ONE warp the 34 rows of a tile
// become a serial loop of that warp, and the whole block waits for it at the end of the pass -- measured: 36 % of the
// block's time in the barrier).  roff[j]: element offset of tile row warp + 8 j inside a plane, formed once per block.
constexpr int PM_RPW = (PM_ROWS + TILE_THREADS / 32 - 1) / (TILE_THREADS / 32);     // rows per warp
__device__ __forceinline__ void pm_issue_tile(float* sm, unsigned bar, const float* plane, const int (&roff)[PM_RPW]) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(PM_TILE_BYTES) : "memory");
    if ((threadIdx.x & 31) == 0) {
        const unsigned dst = pm_smem(sm) + (unsigned)warp * PM_ROW_BYTES;
#pragma unroll
        for (int j = 0; j < PM_RPW; ++j)
            if (warp + 8 * j < PM_ROWS)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst + (unsigned)(8 * j) * PM_ROW_BYTES), "l"(plane + roff[j]), "r"(PM_ROW_BYTES), "r"(bar) : "memory");
    }
}
// border tiles: the same layout through the threads (reflect-pad(1) of the columns, as load_tile)
__device__ __forceinline__ void pm_stage_tile(float* __restrict__ sm, const float* __restrict__ plane, int x0, int y0, int W, int H) {
    const int tx = threadIdx.x % TILE_W, tr = threadIdx.x / TILE_W;
    const int ca = fold_coord(x0 - 1 + tx, W), cb = fold_coord(x0 - 1 + TILE_W + tx, W);
#pragma unroll
    for (int ty = tr; ty < PM_ROWS; ty += TILE_THREADS / TILE_W) {
        const float* row = plane + (long long)fold_coord(y0 - 1 + ty, H) * W;
        sm[ty * PM_PITCH + PM_COL0 + tx] = row[ca];
        if (tx < 2) sm[ty * PM_PITCH + PM_COL0 + TILE_W + tx] = row[cb];
    }
}

template <int C>
__global__ void __launch_bounds__(TILE_THREADS, MD2_PM_MINB) photomin_fwd_bulk_kernel(PmArgs a, float* __restrict__ out, int* __restrict__ argmin) {
    __shared__ __align__(128) float ys[2][PM_ROWS * PM_PITCH];
    __shared__ __align__(128) float xs[2][PM_ROWS * PM_PITCH];
    __shared__ __align__(8) unsigned long long bars[4];          // 0, 1: ys[0], ys[1];  2, 3: xs[0], xs[1]
    const long long HW = (long long)a.W * a.H;
    const int n = blockIdx.z, x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * FWD_TH;
    const int tx = threadIdx.x % TILE_W, tr = threadIdx.x / TILE_W;
    const int gx = x0 + tx, gy0 = y0 + tr * FWD_RPT;
    const float* y = a.target + n * a.target_ns;
    const int S = a.S;
    // block-uniform: can the copy engine stage this tile (16-byte aligned rows, no reflected columns)?
    bool bulk = (a.W & 3) == 0 && x0 >= 4 && x0 + TILE_W + 4 <= a.W && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    const float* pb[MAX_S];                                      // (registers: no run-time indexing of the parameter block)
#pragma unroll
    for (int s = 0; s < MAX_S; ++s) {
        pb[s] = s < S ? a.pred[s] + n * a.pred_ns[s] : y;
        bulk = bulk && (reinterpret_cast<uintptr_t>(pb[s]) & 15) == 0;
    }
    auto pred_of = [&](int s) {
        const float* q = pb[0];
#pragma unroll
        for (int j = 1; j < MAX_S; ++j) q = (s == j) ? pb[j] : q;
        return q;
    };
    const unsigned bar0 = pm_smem(&bars[0]);
    int roff[PM_RPW];
#pragma unroll
    for (int j = 0; j < PM_RPW; ++j) roff[j] = fold_coord(y0 - 1 + (int)(threadIdx.x >> 5) + 8 * j, a.H) * a.W + (x0 - 1 - PM_COL0);
    if (bulk) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * j) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        pm_issue_tile(ys[0], bar0, y, roff);
        pm_issue_tile(xs[0], bar0 + 16u, pb[0], roff);
    }
    float ss[MAX_S][FWD_RPT], l1[MAX_S][FWD_RPT];
#pragma unroll
    for (int s = 0; s < MAX_S; ++s)
#pragma unroll
        for (int r = 0; r < FWD_RPT; ++r) { ss[s][r] = 0.f; l1[s][r] = 0.f; }
    const int passes = C * S;
#pragma unroll 1
    for (int k = 0, c = 0, s = 0; k < passes; ++k) {
        const float* ysm = ys[c & 1];
        const float* xsm = xs[k & 1];
        if (bulk) {
            // prefetch the next pass: its buffers were last read in pass k - 1, which every thread has left (barrier below)
            if (k + 1 < passes) {
                const int s1 = (s + 1 == S) ? 0 : s + 1, c1 = (s + 1 == S) ? c + 1 : c;
                if (s1 == 0) pm_issue_tile(ys[c1 & 1], bar0 + 8u * (c1 & 1), y + c1 * HW, roff);
                pm_issue_tile(xs[(k + 1) & 1], bar0 + 16u + 8u * ((k + 1) & 1), pred_of(s1) + c1 * HW, roff);
            }
            if (s == 0) pm_bar_wait(bar0 + 8u * (c & 1), (c >> 1) & 1);
            pm_bar_wait(bar0 + 16u + 8u * (k & 1), (k >> 1) & 1);
        } else {
            if (s == 0) pm_stage_tile(ys[c & 1], y + c * HW, x0, y0, a.W, a.H);
            pm_stage_tile(xs[k & 1], pred_of(s) + c * HW, x0, y0, a.W, a.H);
            __syncthreads();
        }
        {
            const int cbase = (tr * FWD_RPT + 1) * PM_PITCH + PM_COL0 + 1 + tx;
            const float xr = xsm[cbase], yr = ysm[cbase];
            float ha[3], hb[3], haa[3], hbb[3], hab[3], yc[3], xc[3];
            float accs[FWD_RPT], accl[FWD_RPT];
#pragma unroll
            for (int j = 0; j < FWD_RPT + 2; ++j) {
                const int o = (tr * FWD_RPT + j) * PM_PITCH + PM_COL0 + tx;
                const float al = xsm[o] - xr, ac = xsm[o + 1] - xr, ar = xsm[o + 2] - xr;
                const float bl = ysm[o] - yr, bc = ysm[o + 1] - yr, br = ysm[o + 2] - yr;
                const int m3 = j % 3;
                ha[m3] = al + ac + ar; hb[m3] = bl + bc + br;
                haa[m3] = fmaf(ar, ar, fmaf(ac, ac, al * al)); hbb[m3] = fmaf(br, br, fmaf(bc, bc, bl * bl));
                hab[m3] = fmaf(ar, br, fmaf(ac, bc, al * bl));
                xc[m3] = xsm[o + 1]; yc[m3] = ysm[o + 1];
                if (j >= 2) {
                    const int m = (j - 1) % 3;
                    accs[j - 2] = ssim_from_sums(xr, yr, ha[0] + ha[1] + ha[2], hb[0] + hb[1] + hb[2], haa[0] + haa[1] + haa[2],
                                                 hbb[0] + hbb[1] + hbb[2], hab[0] + hab[1] + hab[2]);
                    accl[j - 2] = fabsf(yc[m] - xc[m]);
                }
            }
#pragma unroll
            for (int sq = 0; sq < MAX_S; ++sq)      // (static register indexing of the per-source accumulators)
                if (sq == s) {
#pragma unroll
                    for (int r = 0; r < FWD_RPT; ++r) { ss[sq][r] += accs[r]; l1[sq][r] += accl[r]; }
                }
        }
        __syncthreads();                            // every thread is done with the buffers of pass k
        if (++s == S) { s = 0; ++c; }
    }
    if (gx >= a.W) return;
#pragma unroll
    for (int r = 0; r < FWD_RPT; ++r) {
        const int gy = gy0 + r;
        if (gy >= a.H) break;
        const long long i = (long long)n * HW + (long long)gy * a.W + gx;
        float best = 0.f; int bi = -1;
        if (a.mask) best = a.mask[i];
#pragma unroll
        for (int sq = 0; sq < MAX_S; ++sq) {
            if (sq >= S) break;
            const float pe = a.alpha * (ss[sq][r] * (1.0f / C)) + (1.0f - a.alpha) * (l1[sq][r] * (1.0f / C));
            if ((sq == 0 && !a.mask) || pe < best) { best = pe; bi = sq; }
        }
        out[i] = best;
        if (argmin) argmin[i] = bi;
    }
}

template <int C>
static void launch_photomin_fwd(const PmArgs& a, float* out, int* argmin, dim3 g, cudaStream_t st) {
    // (with fewer than four (channel, source) passes there is nothing to overlap the copies' latency with: measured 12 us
    // through the threads against 16 us at 416x128x8, C = 1)
    if (MD2_PM_BULK && C * a.S >= 4) photomin_fwd_bulk_kernel<C><<<g, TILE_THREADS, 0, st>>>(a, out, argmin);
    else photomin_fwd_kernel<C><<<g, TILE_THREADS, 0, st>>>(a, out, argmin);
}

// Backward, two phases per tile and channel like ssim_bwd_kernel; a window's coefficients are those of ITS selected
// source (argmin of the forward), so the tiles of all S sources are staged side by side.
template <int C>
__global__ void __launch_bounds__(TILE_THREADS) photomin_bwd_kernel(PmArgs a, const float* __restrict__ gout, const int* __restrict__ argmin,
                                                                    float* __restrict__ gtarget, float* __restrict__ gmask) {
    constexpr int SW = TILE_W + 4, SH = TILE_H + 4, CW = TILE_W + 2, CH = TILE_H + 2;
    __shared__ float ys[SW * SH], xs[MAX_S][SW * SH];
    __shared__ float cf[5][CW * CH];
    __shared__ int sel[CW * CH];
    const long long HW = (long long)a.W * a.H;
    const int n = blockIdx.z, x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    const int tx = threadIdx.x % TILE_W, ty = threadIdx.x / TILE_W;
    const int jx = x0 + tx, jy = y0 + ty;
    const bool in = jx < a.W && jy < a.H;
    const float* y = a.target + n * a.target_ns;
    const int* am = argmin ? argmin + (long long)n * HW : nullptr;
    const float ks = a.alpha * (1.0f / C) * (-0.5f), kl = (1.0f - a.alpha) * (1.0f / C);
    for (int k = threadIdx.x; k < CW * CH; k += TILE_THREADS) {      // selected source of every window around the tile (-1: mask / outside)
        const int wy = k / CW, wx = k - wy * CW;
        const int qx = x0 - 1 + wx, qy = y0 - 1 + wy;
        sel[k] = (qx >= 0 && qx < a.W && qy >= 0 && qy < a.H) ? (am ? am[qy * a.W + qx] : 0) : -1;
    }
    const long long i = (long long)n * HW + (long long)(in ? jy : 0) * a.W + (in ? jx : 0);
    const int sj = in ? (am ? am[(long long)jy * a.W + jx] : 0) : -1;
    const float gj = in ? gout[i] : 0.f;
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
        __syncthreads();
        load_tile<2>(ys, y + c * HW, x0, y0, a.W, a.H);
#pragma unroll
        for (int s = 0; s < MAX_S; ++s)
            if (s < a.S) load_tile<2>(xs[s], a.pred[s] + n * a.pred_ns[s] + c * HW, x0, y0, a.W, a.H);
        __syncthreads();
        for (int k = threadIdx.x; k < CW * CH; k += TILE_THREADS) {
            const int wy = k / CW, wx = k - wy * CW;
            const int sq = sel[k];
            float cc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (sq >= 0) {
                float xv[9], yv[9];
                window_samples<SW>(xs[sq], wx + 1, wy + 1, xv);
                window_samples<SW>(ys, wx + 1, wy + 1, yv);
                const WinXY w = window_core(xv, yv);
                const int qx = x0 - 1 + wx, qy = y0 - 1 + wy;
                const float kk = gout[(long long)n * HW + (long long)qy * a.W + qx] * ks * w.pass;
                cc[0] = kk * w.ax; cc[1] = kk * w.bx; cc[2] = kk * w.g; cc[3] = kk * w.ay; cc[4] = kk * w.by;
            }
#pragma unroll
            for (int j = 0; j < 5; ++j) cf[j][k] = cc[j];
        }
        __syncthreads();
        if (in) {
            const float yj = ys[(ty + 2) * SW + tx + 2];
            float gp[MAX_S], gt = 0.f;
#pragma unroll
            for (int s = 0; s < MAX_S; ++s) gp[s] = 0.f;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int k = (ty + 1 + dy) * CW + tx + 1 + dx;
                    const int sq = sel[k];
                    if (sq < 0) continue;
                    const float f = fold_w(jx, dx, a.W) * fold_w(jy, dy, a.H);
                    const float xj = xs[sq][(ty + 2) * SW + tx + 2];
                    const float vx = f * (cf[0][k] + cf[1][k] * xj + cf[2][k] * yj);
#pragma unroll
                    for (int s = 0; s < MAX_S; ++s) if (s == sq) gp[s] += vx;
                    gt = fmaf(f, cf[3][k] + cf[4][k] * yj + cf[2][k] * xj, gt);
                }
            if (sj >= 0) {
                const float df = xs[sj][(ty + 2) * SW + tx + 2] - yj;
                const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
#pragma unroll
                for (int s = 0; s < MAX_S; ++s) if (s == sj) gp[s] += gj * kl * sg;
                gt -= gj * kl * sg;
            }
            const long long o = ((long long)n * C + c) * HW + (long long)jy * a.W + jx;
#pragma unroll
            for (int s = 0; s < MAX_S; ++s)
                if (s < a.S && a.gpred[s]) a.gpred[s][o] = gp[s];
            if (gtarget) gtarget[o] = gt;
        }
    }
    if (in && gmask) gmask[i] = (sj < 0) ? gj : 0.f;
}

int launch_automask(md2_ctx* ctx, int S, const float* const* frames, const int64_t* frame_ns, const float* target, int64_t target_ns,
                    float* out, int W, int H, int C, int N, cudaStream_t st) {
    PmArgs a;
    memset(&a, 0, sizeof(a));
    a.S = S; a.W = W; a.H = H; a.C = C; a.N = N;
    for (int s = 0; s < S; ++s) { a.pred[s] = frames[s]; a.pred_ns[s] = frame_ns[s]; }
    a.target = target; a.target_ns = target_ns; a.mask = nullptr; a.alpha = PHOTO_ALPHA;
    MD2_REQUIRE(N <= 65535 && cdiv(H, FWD_TH) <= 65535, "N and H / 32 must be <= 65535");
    const dim3 g(cdiv(W, TILE_W), cdiv(H, FWD_TH), N);
    if (C == 1) launch_photomin_fwd<1>(a, out, nullptr, g, st);
    else launch_photomin_fwd<3>(a, out, nullptr, g, st);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------
// A8 smooth_loss (src/utils.jl:143-173), optional mean-normalisation (src/training.jl:64-65)
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) smooth_stats_kernel(const float* __restrict__ disp, const float* __restrict__ img,
                                                           long long img_ns, float* __restrict__ partial, int W, int H) {
    __shared__ float scratch[3 * 8];
    const int n = blockIdx.y;
    const long long HW = (long long)W * H;
    float v[3] = {0.f, 0.f, 0.f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x)
        stats_pixel<C>(disp + n * HW, img + n * img_ns, i, W, H, v[0], v[1], v[2]);
    block_sum<3>(v, scratch);
    if (threadIdx.x == 0) {
        float* o = partial + ((long long)n * gridDim.x + blockIdx.x) * NSTAT;
        o[0] = 0.f; o[1] = v[0]; o[2] = v[1]; o[3] = v[2];
    }
}

__global__ void smooth_final_kernel(const float* __restrict__ stats, float* __restrict__ out, int W, int H, int N, int normalize) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const float one = 1.0f;
        *out = loss_from_stats(stats, W, H, N, 1, &one, 1.0f, normalize);
    }
}

template <int C>
__global__ void smooth_bwd_kernel(const float* __restrict__ disp, const float* __restrict__ img, long long img_ns,
                                  const float* __restrict__ stats, float gout, float* __restrict__ gdisp,
                                  float* __restrict__ gimg, int W, int H, int N, int normalize) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long HW = (long long)W * H;
    if (i >= HW * N) return;
    const int n = (int)(i / HW), p = (int)(i % HW), x = p % W, y = p / W;
    const float* d = disp + n * HW;
    const float* t = img + n * img_ns;
    const float cx = 1.0f / ((float)(W - 1) * (float)H * (float)N), cy = 1.0f / ((float)W * (float)(H - 1) * (float)N);
    float sA = gout, sB = 0.f;
    if (normalize) {
        const float* st = stats + (long long)n * NSTAT;
        const float m = st[3] / (float)HW + 1e-7f;
        sA = gout / m;
        sB = gout * (cx * st[1] + cy * st[2]) / (m * m * (float)HW);
    }
    float gh = 0.f;
    float gi[C];
#pragma unroll
    for (int c = 0; c < C; ++c) gi[c] = 0.f;
    // the four pairs this pixel takes part in: (p,p+1), (p-1,p), (p,p+W), (p-W,p)
    const int offs[4] = {1, -1, W, -W};
    const bool ok[4] = {x + 1 < W, x > 0, y + 1 < H, y > 0};
    const float cw[4] = {cx, cx, cy, cy};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!ok[k]) continue;
        const int a = (k & 1) ? p + offs[k] : p;       // first pixel of the pair
        const int b = (k & 1) ? p : p + offs[k];       // second pixel
        float g = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) g += fabsf(t[c * HW + a] - t[c * HW + b]);
        const float e = expf(-g * (1.0f / C));
        const float dd = d[a] - d[b];
        const float sg = dd > 0.f ? 1.f : (dd < 0.f ? -1.f : 0.f);
        const float side = (k & 1) ? -1.f : 1.f;       // p is the second pixel for odd k
        gh += side * cw[k] * sg * e;
        if (gimg) {
            const float term = sA * cw[k] * fabsf(dd) * e * (1.0f / C);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float dt = t[c * HW + a] - t[c * HW + b];
                const float st = dt > 0.f ? 1.f : (dt < 0.f ? -1.f : 0.f);
                gi[c] -= side * term * st;
            }
        }
    }
    gdisp[i] = sA * gh - sB;
    if (gimg)
#pragma unroll
        for (int c = 0; c < C; ++c) gimg[((long long)n * C + c) * HW + p] = gi[c];
}

}  // namespace md2

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace md2;
#define ST ((cudaStream_t)st)
#define CTX_OK() MD2_REQUIRE(ctx != nullptr, "null ctx"); MD2_USE_DEVICE(ctx)

static inline void depth_ab(float min_depth, float max_depth, float& a, float& b) {
    const float mind = (float)(1.0 / (double)max_depth), maxd = (float)(1.0 / (double)min_depth);
    a = maxd - mind; b = mind;
}

extern "C" {

int md2_disparity_to_depth_fwd(md2_ctx* ctx, const float* disp, float* depth, int64_t count, float min_depth,
                               float max_depth, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(disp && depth && count >= 0, "bad arguments");
    if (count == 0) return 0;
    float a, b; depth_ab(min_depth, max_depth, a, b);
    d2d_fwd_kernel<<<cdiv(count, 256), 256, 0, ST>>>(disp, depth, count, a, b);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_disparity_to_depth_bwd(md2_ctx* ctx, const float* disp, const float* gdepth, float* gdisp, int64_t count,
                               float min_depth, float max_depth, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(disp && gdepth && gdisp && count >= 0, "bad arguments");
    if (count == 0) return 0;
    float a, b; depth_ab(min_depth, max_depth, a, b);
    d2d_bwd_kernel<<<cdiv(count, 256), 256, 0, ST>>>(disp, gdepth, gdisp, count, a, b);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

int md2_backproject_fwd(md2_ctx* ctx, const float* depth, const float* invK, float* points, int32_t W, int32_t H,
                        int32_t N, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(depth && invK && points && W > 0 && H > 0 && N > 0, "bad arguments");
    backproject_fwd_kernel<<<cdiv((long long)W * H * N, 256), 256, 0, ST>>>(depth, invK, points, W, H, N);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_backproject_bwd(md2_ctx* ctx, const float* gpoints, const float* invK, float* gdepth, int32_t W, int32_t H,
                        int32_t N, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(gpoints && invK && gdepth && W > 0 && H > 0 && N > 0, "bad arguments");
    backproject_bwd_kernel<<<cdiv((long long)W * H * N, 256), 256, 0, ST>>>(gpoints, invK, gdepth, W, H, N);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

int md2_project_fwd(md2_ctx* ctx, const float* points, const float* K, const float* R, const float* t, float* uv,
                    int32_t W, int32_t H, int32_t N, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(points && K && R && t && uv && W > 1 && H > 1 && N > 0, "bad arguments");
    project_fwd_kernel<<<cdiv((long long)W * H * N, 256), 256, 0, ST>>>(points, K, R, t, uv, W, H, N);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_project_bwd(md2_ctx* ctx, const float* points, const float* K, const float* R, const float* t,
                    const float* guv, float* gpoints, float* gR, float* gt, int32_t W, int32_t H, int32_t N,
                    md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(points && K && R && t && guv && gpoints && gR && gt && W > 1 && H > 1 && N > 0, "bad arguments");
    const int bpi = max(1, min(128, cdiv((long long)W * H, 1024)));
    float* partial = (float*)ws_get(ctx, MD2_WS_PARTIAL, sizeof(float) * (size_t)bpi * N * 12);
    if (!partial) return 1;
    project_bwd_kernel<<<dim3(bpi, N), 256, 0, ST>>>(points, K, R, t, guv, gpoints, partial, W, H, N);
    MD2_LAUNCH_CHECK(ctx);
    return launch_reduce_partials(ctx, partial, gR, 9, gt, N, bpi, 12, ST);
}

int md2_so3_exp_map_fwd(md2_ctx* ctx, const float* rvec, float* R, int32_t N, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(rvec && R && N > 0, "bad arguments");
    so3_fwd_kernel<<<cdiv(N, 64), 64, 0, ST>>>(rvec, R, N);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_so3_exp_map_bwd(md2_ctx* ctx, const float* rvec, const float* gR, float* grvec, int32_t N, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(rvec && gR && grvec && N > 0, "bad arguments");
    so3_bwd_kernel<<<cdiv(N, 64), 64, 0, ST>>>(rvec, gR, grvec, N);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_hat_fwd(md2_ctx* ctx, const float* rvec, float* S, int32_t N, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(rvec && S && N > 0, "bad arguments");
    hat_fwd_kernel<<<cdiv(N, 64), 64, 0, ST>>>(rvec, S, N);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_hat_bwd(md2_ctx* ctx, const float* gS, float* grvec, int32_t N, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(gS && grvec && N > 0, "bad arguments");
    hat_bwd_kernel<<<cdiv(N, 64), 64, 0, ST>>>(gS, grvec, N);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_compose_T_fwd(md2_ctx* ctx, const float* rvec, const float* tvec, int32_t invert, float* R, float* t,
                      int32_t N, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(rvec && tvec && R && t && N > 0, "bad arguments");
    compose_fwd_kernel<<<cdiv(N, 64), 64, 0, ST>>>(rvec, tvec, invert, R, t, N);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_compose_T_bwd(md2_ctx* ctx, const float* rvec, const float* tvec, int32_t invert, const float* gR,
                      const float* gt, float* grvec, float* gtvec, int32_t N, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(rvec && tvec && grvec && gtvec && N > 0, "bad arguments");
    compose_bwd_kernel<<<cdiv(N, 64), 64, 0, ST>>>(rvec, tvec, invert, gR, gt, grvec, gtvec, N);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

int md2_grid_sample_fwd(md2_ctx* ctx, const float* input, const float* grid, float* out, int32_t W, int32_t H,
                        int32_t C, int32_t N, int32_t Wo, int32_t Ho, int32_t padding_mode, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(input && grid && out && W > 0 && H > 0 && C > 0 && N > 0 && Wo > 0 && Ho > 0, "bad arguments");
    MD2_REQUIRE(padding_mode == 0 || padding_mode == 1, "padding_mode must be 0 (:zeros) or 1 (:border)");
    grid_sample_fwd_kernel<<<cdiv((long long)Wo * Ho * N, 256), 256, 0, ST>>>(input, grid, out, W, H, C, N, Wo, Ho, padding_mode);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_grid_sample_bwd(md2_ctx* ctx, const float* input, const float* grid, const float* gout, float* ginput,
                        float* ggrid, int32_t W, int32_t H, int32_t C, int32_t N, int32_t Wo, int32_t Ho,
                        int32_t padding_mode, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(input && grid && gout && W > 0 && H > 0 && C > 0 && N > 0 && Wo > 0 && Ho > 0, "bad arguments");
    MD2_REQUIRE(padding_mode == 0 || padding_mode == 1, "padding_mode must be 0 (:zeros) or 1 (:border)");
    grid_sample_bwd_kernel<<<cdiv((long long)Wo * Ho * N, 256), 256, 0, ST>>>(input, grid, gout, ginput, ggrid, W, H, C, N, Wo, Ho, padding_mode);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

int md2_ssim_fwd(md2_ctx* ctx, const float* x, const float* y, float* out, int32_t W, int32_t H, int32_t C, int32_t N,
                 md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(x && y && out && W >= 2 && H >= 2 && C > 0 && N > 0, "bad arguments (W,H >= 2)");
    const long long planes = (long long)C * N;
    MD2_REQUIRE(planes <= 65535 && cdiv(H, TILE_H) <= 65535, "C * N and H / 8 must be <= 65535");
    ssim_fwd_kernel<<<dim3(cdiv(W, TILE_W), cdiv(H, TILE_H), (unsigned)planes), TILE_THREADS, 0, ST>>>(x, y, out, W, H);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_ssim_bwd(md2_ctx* ctx, const float* x, const float* y, const float* gout, float* gx, float* gy, int32_t W,
                 int32_t H, int32_t C, int32_t N, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(x && y && gout && W >= 2 && H >= 2 && C > 0 && N > 0, "bad arguments (W,H >= 2)");
    const long long planes = (long long)C * N;
    MD2_REQUIRE(planes <= 65535 && cdiv(H, TILE_H) <= 65535, "C * N and H / 8 must be <= 65535");
    ssim_bwd_kernel<<<dim3(cdiv(W, TILE_W), cdiv(H, TILE_H), (unsigned)planes), TILE_THREADS, 0, ST>>>(x, y, gout, gx, gy, W, H);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

static int fill_pm(PmArgs& a, int32_t S, const float* const* pred, const int64_t* pred_ns, const float* target,
                   int64_t target_ns, const float* mask, float alpha, int W, int H, int C, int N) {
    MD2_REQUIRE(S >= 1 && S <= MAX_S, "S must be 1 or 2");
    MD2_REQUIRE(C == 1 || C == 3, "C must be 1 or 3");
    MD2_REQUIRE(pred && pred_ns && target && W >= 2 && H >= 2 && N > 0, "bad arguments (W,H >= 2)");
    a.S = S; a.W = W; a.H = H; a.C = C; a.N = N;
    for (int s = 0; s < MAX_S; ++s) {
        a.pred[s] = s < S ? pred[s] : nullptr;
        a.pred_ns[s] = s < S ? pred_ns[s] : 0;
        a.gpred[s] = nullptr;
        if (s < S) MD2_REQUIRE(pred[s] != nullptr, "null prediction");
    }
    a.target = target; a.target_ns = target_ns; a.mask = mask; a.alpha = alpha;
    return 0;
}

int md2_photometric_min_fwd(md2_ctx* ctx, int32_t S, const float* const* pred, const int64_t* pred_image_stride,
                            const float* target, int64_t target_image_stride, const float* mask, float alpha,
                            float* out, int32_t* argmin, int32_t W, int32_t H, int32_t C, int32_t N, md2_stream st) {
    CTX_OK();
    PmArgs a;
    if (fill_pm(a, S, pred, pred_image_stride, target, target_image_stride, mask, alpha, W, H, C, N)) return 1;
    MD2_REQUIRE(out != nullptr, "null output");
    MD2_REQUIRE(N <= 65535 && cdiv(H, FWD_TH) <= 65535, "N and H / 32 must be <= 65535");
    const dim3 g(cdiv(W, TILE_W), cdiv(H, FWD_TH), N);
    if (C == 1) launch_photomin_fwd<1>(a, out, argmin, g, ST);
    else launch_photomin_fwd<3>(a, out, argmin, g, ST);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

int md2_photometric_min_bwd(md2_ctx* ctx, int32_t S, const float* const* pred, const int64_t* pred_image_stride,
                            const float* target, int64_t target_image_stride, const float* mask, float alpha,
                            const float* gout, const int32_t* argmin, float* const* gpred, float* gtarget,
                            float* gmask, int32_t W, int32_t H, int32_t C, int32_t N, md2_stream st) {
    CTX_OK();
    PmArgs a;
    if (fill_pm(a, S, pred, pred_image_stride, target, target_image_stride, mask, alpha, W, H, C, N)) return 1;
    MD2_REQUIRE(gout != nullptr, "null upstream gradient");
    MD2_REQUIRE(argmin != nullptr || (S == 1 && !mask), "argmin from the forward call is required when S > 1 or a mask is used");
    for (int s = 0; s < S; ++s) a.gpred[s] = gpred ? gpred[s] : nullptr;
    MD2_REQUIRE(N <= 65535 && cdiv(H, TILE_H) <= 65535, "N and H / 8 must be <= 65535");
    const dim3 g(cdiv(W, TILE_W), cdiv(H, TILE_H), N);
    if (C == 1) photomin_bwd_kernel<1><<<g, TILE_THREADS, 0, ST>>>(a, gout, argmin, gtarget, gmask);
    else photomin_bwd_kernel<3><<<g, TILE_THREADS, 0, ST>>>(a, gout, argmin, gtarget, gmask);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

static int smooth_stats(md2_ctx* ctx, const float* disp, const float* image, int64_t image_stride, int W, int H,
                        int C, int N, float** stats_out, cudaStream_t st) {
    const int bpi = max(1, min(64, cdiv((long long)W * H, 2048)));
    float* partial = (float*)ws_get(ctx, MD2_WS_MISC, sizeof(float) * (size_t)bpi * N * NSTAT);
    float* stats = (float*)ws_get(ctx, MD2_WS_STATS, sizeof(float) * (size_t)N * NSTAT);
    if (!partial || !stats) return 1;
    if (C == 1) smooth_stats_kernel<1><<<dim3(bpi, N), 256, 0, st>>>(disp, image, image_stride, partial, W, H);
    else smooth_stats_kernel<3><<<dim3(bpi, N), 256, 0, st>>>(disp, image, image_stride, partial, W, H);
    MD2_LAUNCH_CHECK(ctx);
    *stats_out = stats;
    return launch_reduce_partials(ctx, partial, stats, NSTAT, nullptr, N, bpi, NSTAT, st);
}

int md2_smooth_loss_fwd(md2_ctx* ctx, const float* disp, const float* image, int64_t image_stride, float* out,
                        int32_t normalize, int32_t W, int32_t H, int32_t C, int32_t N, md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(disp && image && out && W >= 2 && H >= 2 && N > 0, "bad arguments (W,H >= 2)");
    MD2_REQUIRE(C == 1 || C == 3, "C must be 1 or 3");
    float* stats;
    if (smooth_stats(ctx, disp, image, image_stride, W, H, C, N, &stats, ST)) return 1;
    smooth_final_kernel<<<1, 32, 0, ST>>>(stats, out, W, H, N, normalize);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}
int md2_smooth_loss_bwd(md2_ctx* ctx, const float* disp, const float* image, int64_t image_stride, float gout,
                        float* gdisp, float* gimage, int32_t normalize, int32_t W, int32_t H, int32_t C, int32_t N,
                        md2_stream st) {
    CTX_OK();
    MD2_REQUIRE(disp && image && gdisp && W >= 2 && H >= 2 && N > 0, "bad arguments (W,H >= 2)");
    MD2_REQUIRE(C == 1 || C == 3, "C must be 1 or 3");
    float* stats = nullptr;
    if (normalize && smooth_stats(ctx, disp, image, image_stride, W, H, C, N, &stats, ST)) return 1;
    const int g = cdiv((long long)W * H * N, 256);
    if (C == 1) smooth_bwd_kernel<1><<<g, 256, 0, ST>>>(disp, image, image_stride, stats, gout, gdisp, gimage, W, H, N, normalize);
    else smooth_bwd_kernel<3><<<g, 256, 0, ST>>>(disp, image, image_stride, stats, gout, gdisp, gimage, W, H, N, normalize);
    MD2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // extern "C"
