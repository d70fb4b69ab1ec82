/* Plain-C client of the C ABI (include/md2.h): what a non-Python host (the Julia `ccall` binding, a C++ trainer) does.
 * Compiled with gcc against the header and linked to libmd2_b200.so + libcudart (tests/test_gpu_cabi_c.py):
 *
 *   harness <in.bin> <out.bin>
 *
 * in.bin : int32 N, C, H, W, L, then float32 x (N,3,C,H,W), L disparities (N,1,h_l,w_l) with (w_l, h_l) =
 *          (W, H) >> (L-1-l), rvec_0, tvec_0, rvec_1, tvec_1 (N,3 each), K (3,3), invK (3,3) [column-major]
 * out.bin: two result blocks, each { loss, L disparity gradients, grvec_0, gtvec_0, grvec_1, gtvec_1 } as float32:
 *          block 1 from md2_view_synthesis_loss_fwdbwd (device pointers), block 2 from
 *          md2_view_synthesis_loss_fwdbwd_host (host pointers).  Exit code 0 on success; errors go to stderr. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "md2.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } } while (0)
#define MD(call) do { if ((call) != 0) { fprintf(stderr, "%s: %s\n", #call, md2_last_error()); return 3; } } while (0)

static float* rd(FILE* f, size_t n) {
    float* p = (float*)malloc(n * sizeof(float));
    if (!p || fread(p, sizeof(float), n, f) != n) { fprintf(stderr, "short read\n"); exit(4); }
    return p;
}
static float* to_dev(const float* h, size_t n) {
    float* d = NULL;
    if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess || cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { fprintf(stderr, "to_dev failed\n"); exit(5); }
    return d;
}
static float* dev_buf(size_t n) {
    float* d = NULL;
    if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); exit(5); }
    return d;
}
static void wr_dev(FILE* f, const float* d, size_t n) {
    float* h = (float*)malloc(n * sizeof(float));
    if (cudaMemcpy(h, d, n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "D2H failed\n"); exit(6); }
    fwrite(h, sizeof(float), n, f);
    free(h);
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: harness in.bin out.bin\n"); return 1; }
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) { perror(argv[1]); return 1; }
    int32_t hdr[5];
    if (fread(hdr, sizeof(int32_t), 5, fi) != 5) return 4;
    const int N = hdr[0], C = hdr[1], H = hdr[2], W = hdr[3], L = hdr[4];
    if (L < 1 || L > MD2_MAX_SCALES) return 4;
    const size_t img = (size_t)C * H * W, nx = (size_t)N * 3 * img;
    float* x = rd(fi, nx);
    float* disp[MD2_MAX_SCALES]; size_t nd[MD2_MAX_SCALES]; int dw[MD2_MAX_SCALES], dh[MD2_MAX_SCALES];
    for (int l = 0; l < L; ++l) { dw[l] = W >> (L - 1 - l); dh[l] = H >> (L - 1 - l); nd[l] = (size_t)N * dw[l] * dh[l]; disp[l] = rd(fi, nd[l]); }
    float* pose[4];
    for (int k = 0; k < 4; ++k) pose[k] = rd(fi, (size_t)3 * N);
    float* K = rd(fi, 9); float* invK = rd(fi, 9);
    fclose(fi);
    FILE* fo = fopen(argv[2], "wb");
    if (!fo) { perror(argv[2]); return 1; }

    md2_ctx* ctx = NULL;
    MD(md2_create(0, &ctx));
    fprintf(stderr, "%s\n", md2_version());

    /* ---- device-pointer entry point ---- */
    md2_vsl_desc d;
    memset(&d, 0, sizeof(d));
    d.W = W; d.H = H; d.N = N; d.C = C; d.S = 2; d.L = L;
    float* xd = to_dev(x, nx);
    d.target = xd + img; d.target_image_stride = 3 * (int64_t)img;              /* frame 1 of (W,H,C,3,N) */
    d.source[0] = xd; d.source[1] = xd + 2 * img;                                /* frames 0 and 2, as views */
    d.source_image_stride[0] = d.source_image_stride[1] = 3 * (int64_t)img;
    for (int l = 0; l < L; ++l) {
        d.disparity[l] = to_dev(disp[l], nd[l]); d.disp_w[l] = dw[l]; d.disp_h[l] = dh[l];
        d.grad_disparity[l] = dev_buf(nd[l]);
        d.smooth_weight[l] = 1e-3f / (float)(1 << (L - 1 - l));
    }
    d.K = to_dev(K, 9); d.invK = to_dev(invK, 9);
    d.pose_mode = 1;
    d.rot[0] = to_dev(pose[0], 3 * N); d.trans[0] = to_dev(pose[1], 3 * N); d.invert[0] = 1;
    d.rot[1] = to_dev(pose[2], 3 * N); d.trans[1] = to_dev(pose[3], 3 * N); d.invert[1] = 0;
    d.min_depth = 0.1f; d.max_depth = 100.0f; d.loss_scale = 1.0f / (float)L; d.normalize_disparity = 1;
    d.loss = dev_buf(1);
    for (int s = 0; s < 2; ++s) { d.grad_rot[s] = dev_buf(3 * N); d.grad_trans[s] = dev_buf(3 * N); }
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    MD(md2_view_synthesis_loss_fwdbwd(ctx, &d, 1.0f, (md2_stream)st));
    CK(cudaStreamSynchronize(st));
    wr_dev(fo, d.loss, 1);
    for (int l = 0; l < L; ++l) wr_dev(fo, d.grad_disparity[l], nd[l]);
    for (int s = 0; s < 2; ++s) { wr_dev(fo, d.grad_rot[s], 3 * N); wr_dev(fo, d.grad_trans[s], 3 * N); }
    const long long launches = (long long)md2_launch_count(ctx);

    /* ---- host-pointer entry point: the same descriptor with host buffers ---- */
    md2_vsl_desc h = d;
    float loss_h = 0.f;
    h.target = x + img; h.source[0] = x; h.source[1] = x + 2 * img;
    float* gd[MD2_MAX_SCALES]; float* gp[4];
    for (int l = 0; l < L; ++l) { h.disparity[l] = disp[l]; gd[l] = (float*)calloc(nd[l], sizeof(float)); h.grad_disparity[l] = gd[l]; }
    h.K = K; h.invK = invK;
    for (int k = 0; k < 4; ++k) gp[k] = (float*)calloc(3 * N, sizeof(float));
    h.rot[0] = pose[0]; h.trans[0] = pose[1]; h.rot[1] = pose[2]; h.trans[1] = pose[3];
    h.grad_rot[0] = gp[0]; h.grad_trans[0] = gp[1]; h.grad_rot[1] = gp[2]; h.grad_trans[1] = gp[3];
    h.loss = &loss_h;
    for (int rep = 0; rep < 2; ++rep)      /* second call replays the captured CUDA graph */
        MD(md2_view_synthesis_loss_fwdbwd_host(ctx, &h, 1.0f, 2));
    fwrite(&loss_h, sizeof(float), 1, fo);
    for (int l = 0; l < L; ++l) fwrite(gd[l], sizeof(float), nd[l], fo);
    for (int k = 0; k < 4; ++k) fwrite(gp[k], sizeof(float), 3 * N, fo);
    fclose(fo);

    /* error path: a bad descriptor returns non-zero and a message, nothing aborts */
    md2_vsl_desc bad = d;
    bad.C = 2;
    if (md2_view_synthesis_loss_fwdbwd(ctx, &bad, 1.0f, (md2_stream)st) == 0 || strlen(md2_last_error()) == 0) { fprintf(stderr, "bad descriptor was accepted\n"); return 7; }
    MD(md2_destroy(ctx));
    printf("ok launches=%lld loss_dev_path_written loss_host=%.9g\n", launches, (double)loss_h);
    return 0;
}
