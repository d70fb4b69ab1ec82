/* CPU ORACLE no. 2 (test infrastructure, NOT product code).
 *
 * An independent scalar float64 restatement, in plain C loops, of the view-synthesis loss path of pxl-th/Monodepth2.jl
 * with a HAND-DERIVED reverse pass.  It shares no code and no formulation with
 * either oracle/torch_oracle.py (torch primitives + autograd) or the CUDA kernels (marching strips, centred window sums,
 * alpha / beta / gamma box filters, displacement-form projection, Euler-homogeneity form of the mean-normalisation):
 * every adjoint below is the naive chain rule of the forward statement above it.  Two independent restatements that
 * agree to float64 rounding are the strongest oracle available without a Julia toolchain; tests/test_c_oracle.py holds
 * that comparison, the known-answer vectors of the reference's test/runtests.jl for this file, and (-m gpu) the
 * three-way comparison with the CUDA path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may load this library.
 *
 * Parity status: pinned to the known answers of test/runtests.jl (SSIM 0 / 0.49995 / symmetry, smoothness 0.3 /
 * 0.2542299, so3 vs Rodrigues, composeT inverse, depth range, identity warp); the third-party semantics (NNlib
 * grid_sample / upsample_bilinear / pad_reflect / MeanPool, Zygote's minimum / abs / clamp adjoints) are restated from
 * their published behaviour and are "parity unpinned" beyond those vectors.
 *
 * Layout: a Julia (W,H,C,N) array is a C array [n][c][h][w]; x is [n][frame][c][h][w]; pixel coordinates are 1-based
 * (src/utils.jl:47-51); K, invK are 3x3 row-major in the maths convention K[i][j]; poses are [source][n][3].
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define CO_API __attribute__((visibility("default")))

/* ---- src/utils.jl:175-179  disparity_to_depth ------------------------------------------------------------------ */
static double depth_of(double d, double min_depth, double max_depth) {
    const double min_disp = 1.0 / max_depth, max_disp = 1.0 / min_depth;
    return 1.0 / (d * (max_disp - min_disp) + min_disp);
}
CO_API void co_disparity_to_depth(const double* d, long n, double min_depth, double max_depth, double* out) {
    for (long i = 0; i < n; ++i) out[i] = depth_of(d[i], min_depth, max_depth);
}

/* ---- src/utils.jl:101-141  hat / so3_exp_map -------------------------------------------------------------------- */
static void hat3(const double r[3], double Kx[3][3]) {
    memset(Kx, 0, 9 * sizeof(double));
    Kx[1][0] = r[2];  Kx[0][1] = -r[2];
    Kx[2][0] = -r[1]; Kx[0][2] = r[1];
    Kx[2][1] = r[0];  Kx[1][2] = -r[0];
}
static void mat3_mul(const double A[3][3], const double B[3][3], double Cm[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0.0;
            for (int k = 0; k < 3; ++k) a += A[i][k] * B[k][j];
            Cm[i][j] = a;
        }
}
static void so3_exp(const double r[3], double R[3][3]) {
    double Kx[3][3], K2[3][3];
    hat3(r, Kx);
    mat3_mul(Kx, Kx, K2);
    const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    const double thc = th > 1e-4 ? th : 1e-4;                       /* max.(theta, 1e-4), src/utils.jl:110 */
    const double f1 = sin(th) / thc, f2 = (1.0 - cos(th)) / (thc * thc);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = f1 * Kx[i][j] + f2 * K2[i][j] + (i == j ? 1.0 : 0.0);
}
/* pullback of so3_exp: Rb = d loss / d R  ->  rb = d loss / d r (appendix A.4; theta = 0 gives 0/0 = NaN like the
 * reference, README.md:47-49) */
static void so3_exp_bwd(const double r[3], const double Rb[3][3], double rb[3]) {
    double Kx[3][3], K2[3][3];
    hat3(r, Kx);
    mat3_mul(Kx, Kx, K2);
    const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    const int above = th > 1e-4;
    const double thc = above ? th : 1e-4;
    const double s = sin(th), c = cos(th);
    const double f1 = s / thc, f2 = (1.0 - c) / (thc * thc);
    double f1b = 0.0, f2b = 0.0, Kb[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { f1b += Rb[i][j] * Kx[i][j]; f2b += Rb[i][j] * K2[i][j]; }
    /* K2 = K K: Kb = f1 Rb + f2 (Rb K^T + K^T Rb) */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0.0;
            for (int k = 0; k < 3; ++k) a += Rb[i][k] * Kx[j][k] + Kx[k][i] * Rb[k][j];
            Kb[i][j] = f1 * Rb[i][j] + f2 * a;
        }
    /* f1 = sin(th) / thc, f2 = (1 - cos(th)) / thc^2; d thc / d th = [th > 1e-4] */
    const double dthc = above ? 1.0 : 0.0;
    const double df1 = c / thc - dthc * s / (thc * thc);
    const double df2 = s / (thc * thc) - dthc * 2.0 * (1.0 - c) / (thc * thc * thc);
    const double thb = f1b * df1 + f2b * df2;
    /* hat pullback (src/utils.jl:130-141) + theta = sqrt(r . r) */
    const double hp[3] = {Kb[2][1] - Kb[1][2], Kb[0][2] - Kb[2][0], Kb[1][0] - Kb[0][1]};
    for (int i = 0; i < 3; ++i) rb[i] = hp[i] + thb * r[i] / th;
}
CO_API void co_so3_exp_map(const double* rvec, int n, double* R /* [n][3][3] */) {
    for (int i = 0; i < n; ++i) so3_exp(rvec + 3 * i, (double(*)[3])(R + 9 * i));
}
CO_API void co_so3_exp_map_bwd(const double* rvec, int n, const double* Rb, double* rb) {
    for (int i = 0; i < n; ++i) so3_exp_bwd(rvec + 3 * i, (const double(*)[3])(Rb + 9 * i), rb + 3 * i);
}

/* ---- src/utils.jl:181-188  composeT ----------------------------------------------------------------------------- */
static void compose_T(const double r[3], const double t[3], int invert, double Ru[3][3], double tu[3]) {
    double R[3][3];
    so3_exp(r, R);
    if (!invert) {
        memcpy(Ru, R, sizeof(R));
        memcpy(tu, t, 3 * sizeof(double));
        return;
    }
    for (int i = 0; i < 3; ++i) {
        double a = 0.0;
        for (int j = 0; j < 3; ++j) { Ru[i][j] = R[j][i]; a += R[j][i] * (-t[j]); }
        tu[i] = a;
    }
}
/* Rub, tub: cotangents of the (R, t) actually used  ->  rb, tb */
static void compose_T_bwd(const double r[3], const double t[3], int invert, const double Rub[3][3], const double tub[3],
                          double rb[3], double tb[3]) {
    double Rb[3][3];
    if (!invert) {
        memcpy(Rb, Rub, sizeof(Rb));
        memcpy(tb, tub, 3 * sizeof(double));
    } else {
        double R[3][3];
        so3_exp(r, R);
        /* Ru[i][j] = R[j][i];  tu[i] = -sum_j R[j][i] t[j] */
        for (int j = 0; j < 3; ++j) {
            double a = 0.0;
            for (int i = 0; i < 3; ++i) { Rb[j][i] = Rub[i][j] - t[j] * tub[i]; a -= R[j][i] * tub[i]; }
            tb[j] = a;
        }
    }
    so3_exp_bwd(r, (const double(*)[3])Rb, rb);
}
CO_API void co_composeT(const double* rvec, const double* t, int n, int invert, double* R, double* tout) {
    for (int i = 0; i < n; ++i) compose_T(rvec + 3 * i, t + 3 * i, invert, (double(*)[3])(R + 9 * i), tout + 3 * i);
}

/* ---- src/utils.jl:13-39  SSIM: reflect-pad(1), 3x3 stride-1 mean pool -------------------------------------------- */
static int refl(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }   /* pad 1: -1 -> 1, n -> n-2 */
#define SSIM_C1 (0.01 * 0.01)
#define SSIM_C2 (0.03 * 0.03)

typedef struct { double mux, muy, pxx, pyy, pxy; } win_t;
static win_t window(const double* x, const double* y, int H, int W, int h, int w) {
    win_t q = {0, 0, 0, 0, 0};
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int j = refl(h + dy, H) * W + refl(w + dx, W);
            q.mux += x[j]; q.muy += y[j]; q.pxx += x[j] * x[j]; q.pyy += y[j] * y[j]; q.pxy += x[j] * y[j];
        }
    q.mux /= 9.0; q.muy /= 9.0; q.pxx /= 9.0; q.pyy /= 9.0; q.pxy /= 9.0;
    return q;
}
/* raw = (1 - n / d) / 2 before the clamp */
static double ssim_raw(win_t q) {
    const double sx = q.pxx - q.mux * q.mux, sy = q.pyy - q.muy * q.muy, sxy = q.pxy - q.mux * q.muy;
    const double n = (2.0 * q.mux * q.muy + SSIM_C1) * (2.0 * sxy + SSIM_C2);
    const double d = (q.mux * q.mux + q.muy * q.muy + SSIM_C1) * (sx + sy + SSIM_C2);
    return (1.0 - n / d) * 0.5;
}
static double clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }
/* x, y: [planes][H][W] */
CO_API void co_ssim(const double* x, const double* y, int planes, int H, int W, double* out) {
    for (int p = 0; p < planes; ++p)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                out[(p * H + h) * W + w] = clamp01(ssim_raw(window(x + (long)p * H * W, y + (long)p * H * W, H, W, h, w)));
}
/* adjoint of one plane: g[h][w] = d loss / d ssim window value -> xb += d loss / d x (y treated as data) */
/* fc != NULL: forced decisions (word 0 of every pixel, stride fstride): bit fbit says whether the clamp passes */
static void ssim_plane_bwd(const double* x, const double* y, int H, int W, const double* g, double* xb,
                           const int* fc, int fstride, int fbit) {
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
            const double gw = g[h * W + w];
            if (gw == 0.0) continue;
            const win_t q = window(x, y, H, W, h, w);
            const double raw = ssim_raw(q);
            if (fc ? !((fc[(h * W + w) * fstride] >> fbit) & 1) : (raw < 0.0 || raw > 1.0)) continue;   /* clamp passes the gradient on the closed interval */
            const double sx = q.pxx - q.mux * q.mux, sy = q.pyy - q.muy * q.muy, sxy = q.pxy - q.mux * q.muy;
            const double A = 2.0 * q.mux * q.muy + SSIM_C1, B = 2.0 * sxy + SSIM_C2;
            const double Cc = q.mux * q.mux + q.muy * q.muy + SSIM_C1, D = sx + sy + SSIM_C2;
            const double S = A * B / (Cc * D);
            const double Sb = -0.5 * gw;
            const double Ab = Sb * B / (Cc * D), Bb = Sb * A / (Cc * D), Cb = -Sb * S / Cc, Db = -Sb * S / D;
            /* A(mux), B(pxy, mux), Cc(mux), D(pxx, mux) */
            const double muxb = Ab * 2.0 * q.muy + Bb * (-2.0 * q.muy) + Cb * 2.0 * q.mux + Db * (-2.0 * q.mux);
            const double pxxb = Db, pxyb = Bb * 2.0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int j = refl(h + dy, H) * W + refl(w + dx, W);
                    xb[j] += (muxb + 2.0 * x[j] * pxxb + y[j] * pxyb) / 9.0;
                }
        }
}

/* ---- src/utils.jl:159-173  smooth_loss (disparity [n][H][W], image [n][C][H][W]) --------------------------------- */
static double sgn(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0); }
/* returns the loss; if db != NULL adds up * d loss / d disparity */
static double code_sign(int code) { return code == 1 ? 1.0 : (code == 2 ? -1.0 : 0.0); }
/* fc != NULL (N = 1 only): forced signs of the two differences, bits 20-21 / 22-23 of word 0 of every pixel */
static double smooth_loss_fb(const double* d, const double* img, int N, int C, int H, int W, double up, double* db,
                             const int* fc, int fstride) {
    double sx = 0.0, sy = 0.0;
    const double cx = 1.0 / ((double)(W - 1) * H * N), cy = 1.0 / ((double)W * (H - 1) * N);
    for (int n = 0; n < N; ++n) {
        const double* dn = d + (long)n * H * W;
        const double* in = img + (long)n * C * H * W;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const int p = h * W + w;
                if (w + 1 < W) {
                    double gi = 0.0;
                    for (int c = 0; c < C; ++c) gi += fabs(in[c * H * W + p] - in[c * H * W + p + 1]);
                    const double wt = exp(-gi / C), df = dn[p] - dn[p + 1];
                    const double sg = fc ? code_sign((fc[p * fstride] >> 20) & 3) : sgn(df);
                    sx += fabs(df) * wt;
                    if (db) { db[(long)n * H * W + p] += up * cx * sg * wt; db[(long)n * H * W + p + 1] -= up * cx * sg * wt; }
                }
                if (h + 1 < H) {
                    double gi = 0.0;
                    for (int c = 0; c < C; ++c) gi += fabs(in[c * H * W + p] - in[c * H * W + p + W]);
                    const double wt = exp(-gi / C), df = dn[p] - dn[p + W];
                    const double sg = fc ? code_sign((fc[p * fstride] >> 22) & 3) : sgn(df);
                    sy += fabs(df) * wt;
                    if (db) { db[(long)n * H * W + p] += up * cy * sg * wt; db[(long)n * H * W + p + W] -= up * cy * sg * wt; }
                }
            }
    }
    return sx * cx + sy * cy;
}
CO_API double co_smooth_loss(const double* d, const double* img, int N, int C, int H, int W, double* db) {
    return smooth_loss_fb(d, img, N, C, H, W, 1.0, db, NULL, 0);
}

/* ---- [3P] NNlib.upsample_bilinear(x; size), align-corners -------------------------------------------------------- */
typedef struct { int i0, i1; double f; } tap_t;
static tap_t up_tap(int o, int n_in, int n_out) {
    tap_t t;
    const double s = n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 0.0;
    const double c = s * o;
    t.i0 = (int)floor(c);
    if (t.i0 > n_in - 1) t.i0 = n_in - 1;
    t.i1 = t.i0 + 1 < n_in ? t.i0 + 1 : n_in - 1;
    t.f = c - t.i0;
    return t;
}
CO_API void co_upsample_bilinear(const double* in, int planes, int h, int w, int H, int W, double* out) {
    for (int p = 0; p < planes; ++p)
        for (int y = 0; y < H; ++y) {
            const tap_t ty = up_tap(y, h, H);
            for (int x = 0; x < W; ++x) {
                const tap_t tx = up_tap(x, w, W);
                const double* a = in + (long)p * h * w;
                out[((long)p * H + y) * W + x] = (1 - ty.f) * ((1 - tx.f) * a[ty.i0 * w + tx.i0] + tx.f * a[ty.i0 * w + tx.i1]) +
                                                 ty.f * ((1 - tx.f) * a[ty.i1 * w + tx.i0] + tx.f * a[ty.i1 * w + tx.i1]);
            }
        }
}
static void upsample_bilinear_bwd(const double* gout, int planes, int h, int w, int H, int W, double* gin) {
    for (int p = 0; p < planes; ++p)
        for (int y = 0; y < H; ++y) {
            const tap_t ty = up_tap(y, h, H);
            for (int x = 0; x < W; ++x) {
                const tap_t tx = up_tap(x, w, W);
                const double g = gout[((long)p * H + y) * W + x];
                double* a = gin + (long)p * h * w;
                a[ty.i0 * w + tx.i0] += g * (1 - ty.f) * (1 - tx.f);
                a[ty.i0 * w + tx.i1] += g * (1 - ty.f) * tx.f;
                a[ty.i1 * w + tx.i0] += g * ty.f * (1 - tx.f);
                a[ty.i1 * w + tx.i1] += g * ty.f * tx.f;
            }
        }
}

/* ---- [3P] NNlib.grid_sample(input, grid; padding_mode = :border), bilinear, align-corners ------------------------ */
typedef struct { int x0, y0; double fx, fy; int mx, my; } cell_t;
/* g in [-1, 1] -> 0-based coordinate, clipped to the border; *m = 0 where the un-clipped coordinate is <= 0 or >= n-1 */
static double unnorm_clip(double g, int n, int* m) {
    const double c = (g + 1.0) * 0.5 * (n - 1);
    if (!(c > 0.0)) { *m = 0; return 0.0; }
    if (c >= n - 1) { *m = 0; return (double)(n - 1); }
    *m = 1;
    return c;
}
static cell_t cell_of(double gx, double gy, int H, int W) {
    cell_t c;
    const double ix = unnorm_clip(gx, W, &c.mx), iy = unnorm_clip(gy, H, &c.my);
    c.x0 = (int)floor(ix); c.y0 = (int)floor(iy);
    c.fx = ix - c.x0; c.fy = iy - c.y0;
    return c;
}
/* the gather cell and the clip-gradient masks taken from the implementation under test (word 1 + s of the pixel) */
static cell_t cell_forced(double gx, double gy, int H, int W, int word) {
    cell_t c;
    int m;
    const double ix = unnorm_clip(gx, W, &m), iy = unnorm_clip(gy, H, &m);
    c.x0 = word & 0x3fff; c.y0 = (word >> 14) & 0x7fff;
    c.fx = ix - c.x0; c.fy = iy - c.y0;
    c.mx = (word >> 29) & 1; c.my = (word >> 30) & 1;
    return c;
}
static double tap_at(const double* img, int H, int W, int y, int x) { return (x >= 0 && x < W && y >= 0 && y < H) ? img[y * W + x] : 0.0; }
static double sample_plane(const double* img, int H, int W, cell_t c) {
    return tap_at(img, H, W, c.y0, c.x0) * (1 - c.fx) * (1 - c.fy) + tap_at(img, H, W, c.y0, c.x0 + 1) * c.fx * (1 - c.fy) +
           tap_at(img, H, W, c.y0 + 1, c.x0) * (1 - c.fx) * c.fy + tap_at(img, H, W, c.y0 + 1, c.x0 + 1) * c.fx * c.fy;
}
/* input [planes][H][W], grid [H][W][2] (x, y) for one image */
CO_API void co_grid_sample_border(const double* in, const double* grid, int planes, int H, int W, double* out) {
    for (int p = 0; p < H * W; ++p) {
        const cell_t c = cell_of(grid[2 * p], grid[2 * p + 1], H, W);
        for (int k = 0; k < planes; ++k) out[(long)k * H * W + p] = sample_plane(in + (long)k * H * W, H, W, c);
    }
}

/* ---- src/training.jl:21-78, the tail of train_loss after model(...) (and src/simple_depth.jl:25-41 with
 *      normalize_disp = 0, one scale of weight 1) ------------------------------------------------------------------ */
typedef struct {
    int H, W;
    const double *K, *invK;
} dims_t;

/* Forward of the photometric part of one scale for one image n; fills warped [S][C][H][W], cells [S][H][W], cam-space
 * quantities needed by the reverse pass, pe / sel maps.  Returns the sum over the pixels of the selected loss. */
typedef struct {
    double *warped, *pe, *wl;   /* [S][C][HW], [S][HW], [HW] */
    int* sel;                   /* [HW]: source index, -1 = automask */
    cell_t* cells;              /* [S][HW] */
} scratch_t;

static void project_pixel(const dims_t* g, const double Ru[3][3], const double tu[3], double z, int w1, int h1,
                          double X[3], double cam[3], double* gxn, double* gyn) {
    /* Backproject (src/utils.jl:63-65): rays = invK (w, h, 1)^T, points = depth * rays */
    const double pix[3] = {(double)w1, (double)h1, 1.0};
    for (int i = 0; i < 3; ++i) {
        double a = 0.0;
        for (int j = 0; j < 3; ++j) a += g->invK[3 * i + j] * pix[j];
        X[i] = z * a;
    }
    /* Project (src/utils.jl:88-99): K (R X + t), perspective divide with eps 1e-7, normalise to (-1, 1) */
    double Y[3];
    for (int i = 0; i < 3; ++i) Y[i] = Ru[i][0] * X[0] + Ru[i][1] * X[1] + Ru[i][2] * X[2] + tu[i];
    for (int i = 0; i < 3; ++i) cam[i] = g->K[3 * i] * Y[0] + g->K[3 * i + 1] * Y[1] + g->K[3 * i + 2] * Y[2];
    const double q = 1.0 / (cam[2] + 1e-7);
    *gxn = ((cam[0] * q - 1.0) / (g->W - 1.0) - 0.5) * 2.0;
    *gyn = ((cam[1] * q - 1.0) / (g->H - 1.0) - 0.5) * 2.0;
}

CO_API int co_view_synthesis_loss(
    const double* x, int N, int L, int C, int H, int W,                 /* frames [N][L][C][H][W] */
    int n_scales, const double* const* disps, const int* dh, const int* dw,   /* decoder outputs, each [N][1][dh][dw] */
    const double* rvecs, const double* tvecs,                            /* [S][N][3] */
    const double* K, const double* invK, int target_id, int S, const int* source_ids,
    const double* scales, double min_depth, double max_depth, double smooth_w,
    const double* auto_loss,                                             /* [N][H][W] or NULL */
    int normalize_disp, double alpha,
    double* loss_out, double* gx,                                        /* gx [N][L][C][H][W] or NULL (source frames only) */
    double* const* gdisps, double* grvecs, double* gtvecs,               /* gradients (NULL: value only) */
    double* viz_warped, double* viz_loss,                                /* last scale: [S][N][C][H][W], [N][H][W], or NULL */
    const int* choices,   /* NULL, or the discrete decisions of the implementation under test, [n_scales][N][H][W][1+S] as
                           * include/md2.h (md2_vsl_desc.debug_choices) lays them out: the same piece of the piecewise-smooth
                           * loss is then evaluated (values as always; only the branch of every kink is taken from there) */
    int* choices_out)     /* NULL, or receives THIS evaluation's own decisions in the same layout (the gather cell reported as
                           * the kernels keep it: the 2x2 cell inside the image, i.e. x0 <= W-2, y0 <= H-2) */
{
    if (H < 2 || W < 2 || S < 1 || n_scales < 1) return 1;
    const int HW = H * W;
    const int want_grad = gdisps != NULL;
    dims_t g = {H, W, K, invK};
    const double up = 1.0 / n_scales;                                     /* loss / length(scales), src/training.jl:77 */
    double total = 0.0;
    double* D = malloc(sizeof(double) * N * HW);                          /* full-resolution disparity of the scale */
    double* Db = malloc(sizeof(double) * N * HW);
    double* Dhat = malloc(sizeof(double) * N * HW);
    double* Dhatb = malloc(sizeof(double) * N * HW);
    scratch_t sc;
    sc.warped = malloc(sizeof(double) * S * C * HW);
    sc.pe = malloc(sizeof(double) * S * HW);
    sc.wl = malloc(sizeof(double) * HW);
    sc.sel = malloc(sizeof(int) * HW);
    sc.cells = malloc(sizeof(cell_t) * S * HW);
    double* wb = malloc(sizeof(double) * S * C * HW);                     /* d loss / d warped */
    double* gs = malloc(sizeof(double) * HW);
    double (*Ru)[3][3] = malloc(sizeof(double[3][3]) * S * N);
    double (*tu)[3] = malloc(sizeof(double[3]) * S * N);
    double (*Rub)[3][3] = calloc(S * N, sizeof(double[3][3]));
    double (*tub)[3] = calloc(S * N, sizeof(double[3]));
    const double min_disp = 1.0 / max_depth, max_disp = 1.0 / min_depth;

    for (int s = 0; s < S; ++s)
        for (int n = 0; n < N; ++n)   /* Pose / composeT with invert = source_id < target_id (src/training.jl:44-46) */
            compose_T(rvecs + 3 * (s * N + n), tvecs + 3 * (s * N + n), source_ids[s] < target_id, Ru[s * N + n], tu[s * N + n]);

    for (int i = 0; i < n_scales; ++i) {
        /* src/training.jl:49-51: upsample to the input size where needed */
        if (dh[i] == H && dw[i] == W) memcpy(D, disps[i], sizeof(double) * N * HW);
        else co_upsample_bilinear(disps[i], N, dh[i], dw[i], H, W, D);
        memset(Db, 0, sizeof(double) * N * HW);
        double photo = 0.0;
        for (int n = 0; n < N; ++n) {
            const double* tgt = x + ((long)(n * L + target_id) * C) * HW;
            const int fs = 1 + S;
            const int* fc = choices ? choices + ((long)(i * N + n) * HW) * fs : NULL;
            /* ---- warp (src/training.jl:52-57) ---- */
            for (int s = 0; s < S; ++s) {
                const double* src = x + ((long)(n * L + source_ids[s]) * C) * HW;
                for (int h = 0; h < H; ++h)
                    for (int w = 0; w < W; ++w) {
                        const int p = h * W + w;
                        double X[3], cam[3], gxn, gyn;
                        project_pixel(&g, Ru[s * N + n], tu[s * N + n], depth_of(D[n * HW + p], min_depth, max_depth), w + 1, h + 1, X, cam, &gxn, &gyn);
                        const cell_t c = fc ? cell_forced(gxn, gyn, H, W, fc[p * fs + 1 + s]) : cell_of(gxn, gyn, H, W);
                        sc.cells[s * HW + p] = c;
                        for (int ch = 0; ch < C; ++ch) sc.warped[(s * C + ch) * HW + p] = sample_plane(src + (long)ch * HW, H, W, c);
                    }
            }
            /* ---- photometric loss per source (src/training.jl:1-5), min over sources (:14-19), automask (:60-62) ---- */
            for (int s = 0; s < S; ++s)
                for (int h = 0; h < H; ++h)
                    for (int w = 0; w < W; ++w) {
                        const int p = h * W + w;
                        double l1 = 0.0, ss = 0.0;
                        for (int ch = 0; ch < C; ++ch) {
                            const double* xw = sc.warped + (s * C + ch) * HW;
                            l1 += fabs(tgt[ch * HW + p] - xw[p]);
                            ss += clamp01(ssim_raw(window(xw, tgt + (long)ch * HW, H, W, h, w)));
                        }
                        sc.pe[s * HW + p] = alpha * (ss / C) + (1.0 - alpha) * (l1 / C);
                    }
            int* co = choices_out ? choices_out + ((long)(i * N + n) * HW) * fs : NULL;
            if (co) {
                for (int p = 0; p < HW; ++p) {
                    int w0 = 0;
                    for (int s = 0; s < S; ++s) {
                        const cell_t c = sc.cells[s * HW + p];
                        const int x0c = c.x0 > W - 2 ? W - 2 : c.x0, y0c = c.y0 > H - 2 ? H - 2 : c.y0;
                        co[p * fs + 1 + s] = x0c | (y0c << 14) | (c.mx << 29) | (c.my << 30);
                        for (int ch = 0; ch < C; ++ch) {
                            const double* xw = sc.warped + (s * C + ch) * HW;
                            const double raw = ssim_raw(window(xw, tgt + (long)ch * HW, H, W, p / W, p % W));
                            const double df = xw[p] - tgt[ch * HW + p];
                            if (raw >= 0.0 && raw <= 1.0) w0 |= 1 << (2 + s * C + ch);
                            w0 |= (df > 0.0 ? 1 : (df < 0.0 ? 2 : 0)) << (8 + 2 * (s * C + ch));
                        }
                    }
                    const double dx = (p % W) + 1 < W ? D[n * HW + p] - D[n * HW + p + 1] : 0.0;
                    const double dy = (p / W) + 1 < H ? D[n * HW + p] - D[n * HW + p + W] : 0.0;
                    w0 |= (dx > 0.0 ? 1 : (dx < 0.0 ? 2 : 0)) << 20;
                    w0 |= (dy > 0.0 ? 1 : (dy < 0.0 ? 2 : 0)) << 22;
                    co[p * fs] = w0;
                }
            }
            for (int p = 0; p < HW; ++p) {
                double v = sc.pe[p];
                int sel = 0;
                for (int s = 1; s < S; ++s)
                    if (sc.pe[s * HW + p] < v) { v = sc.pe[s * HW + p]; sel = s; }   /* findmin: first index wins ties */
                if (auto_loss && auto_loss[n * HW + p] <= v) { v = auto_loss[n * HW + p]; sel = -1; }   /* the mask is first in the cat */
                if (fc) { sel = (fc[p * fs] & 3) - 1; v = sel < 0 ? auto_loss[n * HW + p] : sc.pe[sel * HW + p]; }
                sc.sel[p] = sel; sc.wl[p] = v;
                if (co) co[p * fs] |= sel + 1;
                photo += v;
            }
            if (i == n_scales - 1) {   /* visualisation outputs (src/training.jl:71-74) */
                if (viz_loss) memcpy(viz_loss + (long)n * HW, sc.wl, sizeof(double) * HW);
                if (viz_warped)
                    for (int s = 0; s < S; ++s) memcpy(viz_warped + ((long)(s * N + n) * C) * HW, sc.warped + (long)s * C * HW, sizeof(double) * C * HW);
            }
            if (!want_grad) continue;
            /* ================= reverse pass of this image's photometric term ================= */
            const double gpe = up / ((double)N * HW);                      /* mean over (W, H, 1, N) */
            memset(wb, 0, sizeof(double) * S * C * HW);
            for (int s = 0; s < S; ++s)
                for (int ch = 0; ch < C; ++ch) {
                    const double* xw = sc.warped + (s * C + ch) * HW;
                    for (int p = 0; p < HW; ++p) {
                        const double gp = sc.sel[p] == s ? gpe : 0.0;
                        gs[p] = gp * alpha / C;
                        const double sg = fc ? code_sign((fc[p * fs] >> (8 + 2 * (s * C + ch))) & 3) : sgn(xw[p] - tgt[ch * HW + p]);
                        wb[(s * C + ch) * HW + p] += gp * (1.0 - alpha) / C * sg;   /* d |T - X| / d X */
                    }
                    ssim_plane_bwd(xw, tgt + (long)ch * HW, H, W, gs, wb + (s * C + ch) * HW, fc, fs, 2 + s * C + ch);
                }
            for (int s = 0; s < S; ++s) {
                const double* src = x + ((long)(n * L + source_ids[s]) * C) * HW;
                double* gsrc = gx ? gx + ((long)(n * L + source_ids[s]) * C) * HW : NULL;
                const double (*R)[3] = Ru[s * N + n];
                for (int h = 0; h < H; ++h)
                    for (int w = 0; w < W; ++w) {
                        const int p = h * W + w;
                        const cell_t c = sc.cells[s * HW + p];
                        double ixb = 0.0, iyb = 0.0;
                        for (int ch = 0; ch < C; ++ch) {
                            const double gw = wb[(s * C + ch) * HW + p];
                            if (gw == 0.0) continue;
                            const double* im = src + (long)ch * HW;
                            const double v00 = tap_at(im, H, W, c.y0, c.x0), v01 = tap_at(im, H, W, c.y0, c.x0 + 1);
                            const double v10 = tap_at(im, H, W, c.y0 + 1, c.x0), v11 = tap_at(im, H, W, c.y0 + 1, c.x0 + 1);
                            ixb += gw * ((v01 - v00) * (1 - c.fy) + (v11 - v10) * c.fy);
                            iyb += gw * ((v10 - v00) * (1 - c.fx) + (v11 - v01) * c.fx);
                            if (gsrc) {
                                double* gi = gsrc + (long)ch * HW;
                                gi[c.y0 * W + c.x0] += gw * (1 - c.fx) * (1 - c.fy);
                                if (c.x0 + 1 < W) gi[c.y0 * W + c.x0 + 1] += gw * c.fx * (1 - c.fy);
                                if (c.y0 + 1 < H) gi[(c.y0 + 1) * W + c.x0] += gw * (1 - c.fx) * c.fy;
                                if (c.x0 + 1 < W && c.y0 + 1 < H) gi[(c.y0 + 1) * W + c.x0 + 1] += gw * c.fx * c.fy;
                            }
                        }
                        /* clip mask, un-normalise ((n-1)/2), normalise (2/(n-1)): back to 1-based pixel coordinates u, v */
                        const double ub = c.mx ? ixb * ((W - 1) * 0.5) * (2.0 / (W - 1.0)) : 0.0;
                        const double vb = c.my ? iyb * ((H - 1) * 0.5) * (2.0 / (H - 1.0)) : 0.0;
                        if (ub == 0.0 && vb == 0.0) continue;
                        const double z = depth_of(D[n * HW + p], min_depth, max_depth);
                        double X[3], cam[3], gxn, gyn;
                        project_pixel(&g, R, tu[s * N + n], z, w + 1, h + 1, X, cam, &gxn, &gyn);
                        const double q = 1.0 / (cam[2] + 1e-7);
                        const double cb[3] = {ub * q, vb * q, -(ub * cam[0] + vb * cam[1]) * q * q};
                        double Yb[3], Xb[3];
                        for (int j = 0; j < 3; ++j) Yb[j] = K[j] * cb[0] + K[3 + j] * cb[1] + K[6 + j] * cb[2];      /* K^T cb */
                        for (int a = 0; a < 3; ++a) {
                            tub[s * N + n][a] += Yb[a];
                            for (int b = 0; b < 3; ++b) Rub[s * N + n][a][b] += Yb[a] * X[b];
                        }
                        for (int b = 0; b < 3; ++b) Xb[b] = R[0][b] * Yb[0] + R[1][b] * Yb[1] + R[2][b] * Yb[2];    /* R^T Yb */
                        double zb = 0.0;
                        for (int b = 0; b < 3; ++b) zb += Xb[b] * (X[b] / z);                                         /* X = z ray */
                        Db[n * HW + p] += -zb * z * z * (max_disp - min_disp);                                       /* z = 1 / (d a + b) */
                    }
            }
        }
        /* ---- smoothness term (src/training.jl:64-68): mean-normalised disparity, weight lambda * scale ---- */
        const double wsm = smooth_w * scales[i];
        double sm = 0.0;
        if (normalize_disp) {
            memset(Dhatb, 0, sizeof(double) * N * HW);
            for (int n = 0; n < N; ++n) {
                double m = 0.0;
                for (int p = 0; p < HW; ++p) m += D[n * HW + p];
                m = m / HW + 1e-7;
                for (int p = 0; p < HW; ++p) Dhat[n * HW + p] = D[n * HW + p] / m;
            }
        }
        {   /* the target frames are not contiguous over n (stride L C H W): evaluate image by image and re-weight */
            for (int n = 0; n < N; ++n) {
                const double* tgt = x + ((long)(n * L + target_id) * C) * HW;
                const double* dsrc = (normalize_disp ? Dhat : D) + (long)n * HW;
                double* dbn = want_grad ? (normalize_disp ? Dhatb : Db) + (long)n * HW : NULL;
                /* one image with N = 1 has means over (W-1) H and W (H-1); the batch means divide by N on top */
                sm += smooth_loss_fb(dsrc, tgt, 1, C, H, W, up * wsm / N, dbn, choices ? choices + ((long)(i * N + n) * HW) * (1 + S) : NULL, 1 + S) / N;
            }
            if (normalize_disp && want_grad)
                for (int n = 0; n < N; ++n) {   /* dhat = d / (mean(d) + eps): chain rule through both occurrences of d */
                    double m = 0.0, dot = 0.0;
                    for (int p = 0; p < HW; ++p) m += D[n * HW + p];
                    m = m / HW + 1e-7;
                    for (int p = 0; p < HW; ++p) dot += Dhatb[n * HW + p] * D[n * HW + p];
                    for (int p = 0; p < HW; ++p) Db[n * HW + p] += Dhatb[n * HW + p] / m - dot / (m * m) / HW;
                }
        }
        total += photo / ((double)N * HW) + sm * wsm;
        if (want_grad) {
            if (dh[i] == H && dw[i] == W)
                for (long p = 0; p < (long)N * HW; ++p) gdisps[i][p] += Db[p];
            else upsample_bilinear_bwd(Db, N, dh[i], dw[i], H, W, gdisps[i]);
        }
    }
    *loss_out = total * up;
    if (want_grad)
        for (int s = 0; s < S; ++s)
            for (int n = 0; n < N; ++n)
                compose_T_bwd(rvecs + 3 * (s * N + n), tvecs + 3 * (s * N + n), source_ids[s] < target_id,
                              (const double(*)[3])Rub[s * N + n], tub[s * N + n], grvecs + 3 * (s * N + n), gtvecs + 3 * (s * N + n));
    free(D); free(Db); free(Dhat); free(Dhatb); free(sc.warped); free(sc.pe); free(sc.wl); free(sc.sel); free(sc.cells);
    free(wb); free(gs); free(Ru); free(tu); free(Rub); free(tub);
    return 0;
}
