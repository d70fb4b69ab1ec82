"""CPU ORACLE (test infrastructure, NOT product code).

Op-for-op PyTorch-CPU restatement of the view-synthesis loss path of
pxl-th/Monodepth2.jl, with torch autograd standing in for Zygote.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
arm may import this file; the product package never does.

Parity status: pinned against every known-answer vector of the reference's own
test-suite (test/runtests.jl, see tests/test_oracle_golden.py).  The semantics
of the third-party primitives the reference calls (NNlib grid_sample /
upsample_bilinear / pad_reflect / MeanPool, Zygote adjoints of minimum / abs /
clamp) are NOT under /root/reference and Julia is not available in this image,
so for those the oracle is "parity unpinned" beyond the reference's tests: they
are restated with the PyTorch primitives NNlib's kernels were ported from.

Memory-layout convention: a Julia column-major array (W,H,C,N) is bit-identical
to a contiguous torch tensor (N,C,H,W); points (3,P,N) == torch (N,P,3);
grid (2,W,H,N) == torch (N,H,W,2).  Small matrices use the natural maths
convention here: K (3,3), R (N,3,3) with R[n,i,j], t (N,3), rvec (N,3).
Pixel coordinates are 1-based exactly like the reference (src/utils.jl:47-51).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# src/utils.jl:175-179  disparity_to_depth
# ----------------------------------------------------------------------------


def disparity_to_depth(disparity, min_depth, max_depth):
    dt = disparity.dtype
    min_disp = torch.tensor(1.0 / max_depth, dtype=dt)
    max_disp = torch.tensor(1.0 / min_depth, dtype=dt)
    return 1.0 / (disparity * (max_disp - min_disp) + min_disp)


# ----------------------------------------------------------------------------
# src/utils.jl:13-39  SSIM (reflect-pad 1, 3x3 stride-1 mean pool)
# ----------------------------------------------------------------------------


class SSIM:
    def __init__(self):
        self.c1 = 0.01 ** 2
        self.c2 = 0.03 ** 2

    def __call__(self, x, y):
        # x, y: (N,C,H,W)
        pool = lambda a: F.avg_pool2d(a, 3, 1)
        x_ref = F.pad(x, (1, 1, 1, 1), mode="reflect")
        y_ref = F.pad(y, (1, 1, 1, 1), mode="reflect")
        mu_x, mu_y = pool(x_ref), pool(y_ref)
        c1 = torch.tensor(self.c1, dtype=x.dtype)
        c2 = torch.tensor(self.c2, dtype=x.dtype)
        sigma_x = pool(x_ref * x_ref) - mu_x * mu_x
        sigma_y = pool(y_ref * y_ref) - mu_y * mu_y
        sigma_xy = pool(x_ref * y_ref) - mu_x * mu_y
        ssim_n = (2.0 * mu_x * mu_y + c1) * (2.0 * sigma_xy + c2)
        ssim_d = (mu_x * mu_x + mu_y * mu_y + c1) * (sigma_x + sigma_y + c2)
        return torch.clamp((1.0 - ssim_n / ssim_d) * 0.5, 0.0, 1.0)


# ----------------------------------------------------------------------------
# src/utils.jl:41-65  Backproject
# ----------------------------------------------------------------------------


class Backproject:
    def __init__(self, width, height, dtype=torch.float64):
        w = torch.arange(1, width + 1, dtype=dtype)
        h = torch.arange(1, height + 1, dtype=dtype)
        hh, ww = torch.meshgrid(h, w, indexing="ij")  # p = (h-1)*W + (w-1)
        self.coordinates = torch.stack(
            [ww.reshape(-1), hh.reshape(-1), torch.ones(width * height, dtype=dtype)], 0)  # (3,P)

    def __call__(self, depth, invK):
        # depth (N,P) ; invK (3,3)  ->  points (N,P,3)
        rays = (invK.to(depth.dtype) @ self.coordinates.to(depth.dtype))  # (3,P)
        return depth.unsqueeze(-1) * rays.t().unsqueeze(0)


# ----------------------------------------------------------------------------
# src/utils.jl:67-99  Project + normalize
# ----------------------------------------------------------------------------


class Project:
    def __init__(self, width, height, dtype=torch.float64):
        self.normalizer = torch.tensor([width - 1.0, height - 1.0], dtype=dtype)

    def normalize(self, pixels):
        return (((pixels - 1.0) / self.normalizer.to(pixels.dtype)) - 0.5) * 2.0

    def __call__(self, points, K, R, t):
        # points (N,P,3); K (3,3); R (N,3,3); t (N,3)  ->  (N,P,2) in (-1,1)
        K = K.to(points.dtype)
        cam = torch.einsum("nij,npj->npi", R, points) + t.unsqueeze(1)
        cam = torch.einsum("ij,npj->npi", K, cam)
        denom = 1.0 / (cam[..., 2:3] + 1e-7)
        return self.normalize(cam[..., 0:2] * denom)


# ----------------------------------------------------------------------------
# src/utils.jl:101-141  so3_exp_map / hat ; :181-188 composeT
# ----------------------------------------------------------------------------


def hat(rvec):
    # rvec (N,3) -> (N,3,3)
    N = rvec.shape[0]
    S = torch.zeros(N, 3, 3, dtype=rvec.dtype)
    S[:, 1, 0] = rvec[:, 2]
    S[:, 0, 1] = -rvec[:, 2]
    S[:, 2, 0] = -rvec[:, 1]
    S[:, 0, 2] = rvec[:, 1]
    S[:, 2, 1] = rvec[:, 0]
    S[:, 1, 2] = -rvec[:, 0]
    return S


def hat_pullback(d):
    # src/utils.jl:130-141, d (N,3,3) -> (N,3)
    return torch.stack([d[:, 2, 1] - d[:, 1, 2],
                        -d[:, 2, 0] + d[:, 0, 2],
                        d[:, 1, 0] - d[:, 0, 1]], 1)


def so3_exp_map(rvec):
    N = rvec.shape[0]
    skew = hat(rvec)
    skew2 = skew @ skew
    theta = torch.sqrt((rvec * rvec).sum(1))  # (N,)
    theta_inv = 1.0 / torch.clamp(theta, min=1e-4)
    f1 = (theta_inv * torch.sin(theta)).reshape(N, 1, 1)
    f2 = (theta_inv * theta_inv * (1.0 - torch.cos(theta))).reshape(N, 1, 1)
    return f1 * skew + f2 * skew2 + torch.eye(3, dtype=rvec.dtype)


def composeT(rvec, t, invert):
    # rvec (N,3), t (N,3)
    R = so3_exp_map(rvec)
    if invert:
        R = R.transpose(1, 2)
        t = torch.einsum("nij,nj->ni", R, -t)
    return R, t


# ----------------------------------------------------------------------------
# src/utils.jl:159-173  smooth_loss
# ----------------------------------------------------------------------------


def smooth_loss(disparity, image):
    # disparity (N,H,W) ; image (N,C,H,W)
    ddx = (disparity[:, :, :-1] - disparity[:, :, 1:]).abs()
    ddy = (disparity[:, :-1, :] - disparity[:, 1:, :]).abs()
    idx = (image[:, :, :, :-1] - image[:, :, :, 1:]).abs().mean(1)
    idy = (image[:, :, :-1, :] - image[:, :, 1:, :]).abs().mean(1)
    return (ddx * torch.exp(-idx)).mean() + (ddy * torch.exp(-idy)).mean()


# ----------------------------------------------------------------------------
# src/training.jl:1-19
# ----------------------------------------------------------------------------


def photometric_loss(ssim, predicted, target, alpha=0.85):
    l1 = (target - predicted).abs().mean(1, keepdim=True)
    s = ssim(predicted, target).mean(1, keepdim=True)
    return alpha * s + (1.0 - alpha) * l1


def _min_first(stack):
    # minimum(cat(...; dims=3); dims=3): gradient to the FIRST minimal index
    # (findmin semantics).  torch.min(dim) on CPU returns the first index on
    # ties and routes the gradient to it (amin would split it).
    return stack.min(dim=1, keepdim=True).values


def automasking_loss(ssim, inputs, target, source_ids):
    # inputs (N,L,C,H,W), source_ids 0-based
    return _min_first(torch.cat(
        [photometric_loss(ssim, inputs[:, i], target) for i in source_ids], 1))


def prediction_loss(ssim, predictions, target):
    return _min_first(torch.cat([photometric_loss(ssim, p, target) for p in predictions], 1))


def apply_mask(mask, warp_loss):
    return _min_first(torch.cat([mask, warp_loss], 1))


# ----------------------------------------------------------------------------
# [3P] NNlib primitives restated with the torch ops they were ported from
# ----------------------------------------------------------------------------


def grid_sample(inp, grid, padding_mode="zeros"):
    # NNlib.grid_sample: bilinear, align-corners, padding :zeros (default) or :border
    return F.grid_sample(inp, grid, mode="bilinear", padding_mode=padding_mode, align_corners=True)


def upsample_bilinear(x, size_wh):
    # NNlib.upsample_bilinear(x; size=(W,H)) is align-corners bilinear
    W, H = size_wh
    return F.interpolate(x, size=(H, W), mode="bilinear", align_corners=True)


# ----------------------------------------------------------------------------
# `warp` -- called at src/simple_depth.jl:30-32 but undefined in the reference;
# body inferred from src/training.jl:48-57
# ----------------------------------------------------------------------------


def warp(disp, x, Ps, backproject, project, invK, K, min_depth, max_depth, source_ids):
    # disp (N,1,H,W), x (N,L,C,H,W), Ps list of (R (N,3,3), t (N,3))
    N, _, H, W = disp.shape
    depth = disparity_to_depth(disp, min_depth, max_depth)
    coords = backproject(depth.reshape(N, H * W), invK)
    out = []
    for (R, t), sid in zip(Ps, source_ids):
        uvs = project(coords, K, R, t).reshape(N, H, W, 2)
        out.append(grid_sample(x[:, sid], uvs, padding_mode="border"))
    return out


# ----------------------------------------------------------------------------
# src/training.jl:21-78  train_loss tail (everything after `model(...)`)
# ----------------------------------------------------------------------------


def view_synthesis_loss(x, disparities, rvecs, tvecs, K, invK, *, target_id=1, source_ids=(0, 2),
                        scales=(0.125, 0.25, 0.5, 1.0), min_depth=0.1, max_depth=100.0,
                        disparity_smoothness=1e-3, automasking=False, auto_loss=None,
                        return_viz=False):
    """x (N,L,C,H,W); disparities list of (N,1,h,w); rvecs/tvecs lists of (N,3).
    Returns loss (and, if return_viz, last-scale warped images + warp-loss map)."""
    dt = x.dtype
    N, L, C, H, W = x.shape
    target_x = x[:, target_id]
    ssim = SSIM()
    backproject = Backproject(W, H, dt)
    project = Project(W, H, dt)
    Ps = [composeT(r, t, sid < target_id) for r, t, sid in zip(rvecs, tvecs, source_ids)]
    loss = torch.zeros((), dtype=dt)
    viz = None
    for i, (disparity, scale) in enumerate(zip(disparities, scales)):
        if disparity.shape[-1] != W or disparity.shape[-2] != H:
            disparity = upsample_bilinear(disparity, (W, H))
        warped = warp(disparity, x, Ps, backproject, project, invK, K, min_depth, max_depth,
                      source_ids)
        warp_loss = prediction_loss(ssim, warped, target_x)
        if automasking:
            warp_loss = apply_mask(auto_loss, warp_loss)
        norm_disp = (disparity / (disparity.mean(dim=(2, 3), keepdim=True) + 1e-7))[:, 0]
        disp_loss = smooth_loss(norm_disp, target_x) * disparity_smoothness * scale
        loss = loss + warp_loss.mean() + disp_loss
        if return_viz and i == len(scales) - 1:
            viz = ([w.detach() for w in warped], warp_loss.detach())
    loss = loss / len(scales)
    return (loss, viz) if return_viz else loss


# ----------------------------------------------------------------------------
# src/simple_depth.jl:25-41  objective of the triplet optimiser (config 1)
# ----------------------------------------------------------------------------


def simple_depth_loss(x, disp, rvecs, tvecs, K, invK, *, target_id=1, source_ids=(0, 2),
                      min_depth=0.1, max_depth=100.0):
    """disp (N,1,H,W) raw disparity parameter; un-normalised smoothness, weight 1."""
    dt = x.dtype
    N, L, C, H, W = x.shape
    target_x = x[:, target_id]
    ssim = SSIM()
    Ps = [composeT(r, t, sid < target_id) for r, t, sid in zip(rvecs, tvecs, source_ids)]
    warped = warp(disp, x, Ps, Backproject(W, H, dt), Project(W, H, dt), invK, K,
                  min_depth, max_depth, source_ids)
    warp_loss = prediction_loss(ssim, warped, target_x).mean()
    depth_loss = smooth_loss(disp[:, 0], target_x)
    return warp_loss + depth_loss


# ----------------------------------------------------------------------------
# synthetic inputs: shared by tests and bench so both arms see
# identical data
# ----------------------------------------------------------------------------


def make_K(W, H, f=None, dtype=torch.float32):
    f = 0.58 * W if f is None else f
    K = torch.tensor([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    return K.to(dtype), torch.linalg.inv(K).to(dtype)


def synthetic_batch(N, C, H, W, scales=(0.125, 0.25, 0.5, 1.0), seed=42, full_res_disp=False,
                    pose_sigma=0.01, dtype=torch.float32):
    """Seeded KITTI-shaped triplets: smooth textured target, sub-pixel-to-few-pixel shifted
    sources, low-pass sigmoid disparities at native scale sizes, small non-zero poses."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float64),
                            torch.arange(W, dtype=torch.float64), indexing="ij")
    x = torch.zeros(N, 3, C, H, W, dtype=torch.float64)
    for n in range(N):
        shifts = [(torch.rand(2, generator=g, dtype=torch.float64) - 0.5) * 6.0 for _ in range(3)]
        shifts[1] = torch.zeros(2, dtype=torch.float64)
        for c in range(C):
            k = 6
            fx = (torch.rand(k, generator=g, dtype=torch.float64) - 0.5) * 0.5
            fy = (torch.rand(k, generator=g, dtype=torch.float64) - 0.5) * 0.5
            ph = torch.rand(k, generator=g, dtype=torch.float64) * 6.283
            am = torch.rand(k, generator=g, dtype=torch.float64) / k
            for l in range(3):
                f = torch.zeros(H, W, dtype=torch.float64)
                for j in range(k):
                    f = f + am[j] * torch.sin(fx[j] * (xx + shifts[l][0]) + fy[j] * (yy + shifts[l][1]) + ph[j])
                x[n, l, c] = 0.5 + 0.45 * f
    x = (x + 0.05 * (torch.rand(x.shape, generator=g, dtype=torch.float64) - 0.5)).clamp(0, 1)
    disps = []
    for s in scales:
        h, w = (H, W) if full_res_disp else (max(2, int(round(H * s))), max(2, int(round(W * s))))
        d = torch.randn(N, 1, h, w, generator=g, dtype=torch.float64)
        d = F.avg_pool2d(F.pad(d, (2, 2, 2, 2), mode="replicate"), 5, 1)
        disps.append(torch.sigmoid(2.0 * d).to(dtype))
    rvecs, tvecs = [], []
    for _ in range(2):
        r = pose_sigma * torch.randn(N, 3, generator=g, dtype=torch.float64)
        nr = r.norm(dim=1, keepdim=True)
        r = torch.where(nr < 1e-3, r + 2e-3, r)
        rvecs.append(r.to(dtype))
        tvecs.append((pose_sigma * torch.randn(N, 3, generator=g, dtype=torch.float64)).to(dtype))
    return x.to(dtype), disps, rvecs, tvecs


# ----------------------------------------------------------------------------
# Forced-decision evaluation (test infrastructure for the flip-controlled parity tests).
#
# The loss is only piecewise smooth: the bilinear sampler switches cell at integer coordinates, the border
# clip switches its gradient mask at 1 / size, min-over-sources / automask switch at ties, |.| and the SSIM
# clamp have kinks.  A float32 evaluation lands on either side of such a switch whenever the float64
# quantity is within float32 rounding of it, and both sides are valid sub-gradients.  To compare EVERY
# gradient element strictly, the float64 oracle below takes the discrete decisions from the implementation
# under test (md2.h: md2_vsl_desc.debug_choices) and evaluates the same piece of the piecewise-smooth
# function: values are computed exactly as in view_synthesis_loss, only the branch of each kink is forced.
# ----------------------------------------------------------------------------


def decode_choices(choices, C, S):
    """choices int32 (L,N,H,W,1+S) -> dict of per-scale decisions (see include/md2.h)."""
    w0 = choices[..., 0].long()
    code = lambda v: torch.where(v == 1, 1.0, torch.where(v == 2, -1.0, 0.0)).double()
    out = {"sel": (w0 & 3) - 1,
           "pass": torch.stack([torch.stack([((w0 >> (2 + s * C + c)) & 1).bool() for c in range(C)], 2) for s in range(S)], 2),
           "l1": torch.stack([torch.stack([code((w0 >> (8 + 2 * (s * C + c))) & 3) for c in range(C)], 2) for s in range(S)], 2),
           "smx": code((w0 >> 20) & 3), "smy": code((w0 >> 22) & 3)}   # pass / l1: (L,N,S,C,H,W)
    ws = choices[..., 1:].long()                                       # (L,N,H,W,S)
    out["x0"], out["y0"] = ws & 0x3fff, (ws >> 14) & 0x7fff
    out["mx"], out["my"] = ((ws >> 29) & 1).bool(), ((ws >> 30) & 1).bool()
    return out


def _forced_abs(d, sgn):
    """|d| whose derivative is the given sign (-1, 0, +1)"""
    lin = d * sgn
    return lin + (d.abs() - lin).detach()


def _forced_grid_sample_border(inp, grid, x0, y0, mx, my):
    """bilinear border sampling with the gather cell (x0, y0) and the clip-gradient masks given"""
    N, C, H, W = inp.shape

    def coord(g, size, m):
        i = ((g + 1.0) / 2.0) * (size - 1)
        ic = i + (i.clamp(0, size - 1) - i).detach()       # value of the clip, derivative 1 ...
        return torch.where(m, ic, ic.detach())              # ... where the implementation's mask says so
    fx = (coord(grid[..., 0], W, mx) - x0).unsqueeze(1)
    fy = (coord(grid[..., 1], H, my) - y0).unsqueeze(1)
    flat = inp.reshape(N, C, H * W)
    idx = (y0 * W + x0).reshape(N, 1, H * W).expand(N, C, H * W)
    tap = lambda o: torch.gather(flat, 2, idx + o).reshape(N, C, H, W)
    v00, v01, v10, v11 = tap(0), tap(1), tap(W), tap(W + 1)
    return v00 * (1 - fx) * (1 - fy) + v01 * fx * (1 - fy) + v10 * (1 - fx) * fy + v11 * fx * fy


def _forced_ssim(x, y, passm):
    pool = lambda a: F.avg_pool2d(a, 3, 1)
    x_ref = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y_ref = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x, mu_y = pool(x_ref), pool(y_ref)
    sigma_x = pool(x_ref * x_ref) - mu_x * mu_x
    sigma_y = pool(y_ref * y_ref) - mu_y * mu_y
    sigma_xy = pool(x_ref * y_ref) - mu_x * mu_y
    ssim_n = (2.0 * mu_x * mu_y + 0.01 ** 2) * (2.0 * sigma_xy + 0.03 ** 2)
    ssim_d = (mu_x * mu_x + mu_y * mu_y + 0.01 ** 2) * (sigma_x + sigma_y + 0.03 ** 2)
    raw = (1.0 - ssim_n / ssim_d) * 0.5
    val = raw.clamp(0.0, 1.0)
    return torch.where(passm, raw + (val - raw).detach(), val.detach())


def view_synthesis_loss_forced(x, disparities, rvecs, tvecs, K, invK, choices, *, target_id=1, source_ids=(0, 2),
                               scales=(0.125, 0.25, 0.5, 1.0), min_depth=0.1, max_depth=100.0,
                               disparity_smoothness=1e-3, auto_loss=None, alpha=0.85):
    """view_synthesis_loss with every discrete decision taken from `choices` (int32 (L,N,H,W,1+S))."""
    dt = x.dtype
    N, L, C, H, W = x.shape
    S = len(source_ids)
    ch = decode_choices(choices, C, S)
    target_x = x[:, target_id]
    backproject, project = Backproject(W, H, dt), Project(W, H, dt)
    Ps = [composeT(r, t, sid < target_id) for r, t, sid in zip(rvecs, tvecs, source_ids)]
    loss = torch.zeros((), dtype=dt)
    for i, (disparity, scale) in enumerate(zip(disparities, scales)):
        if disparity.shape[-1] != W or disparity.shape[-2] != H:
            disparity = upsample_bilinear(disparity, (W, H))
        depth = disparity_to_depth(disparity, min_depth, max_depth)
        coords = backproject(depth.reshape(N, H * W), invK)
        pes = []
        for s, ((R, t), sid) in enumerate(zip(Ps, source_ids)):
            uvs = project(coords, K, R, t).reshape(N, H, W, 2)
            warped = _forced_grid_sample_border(x[:, sid], uvs, ch["x0"][i, ..., s], ch["y0"][i, ..., s],
                                                ch["mx"][i, ..., s], ch["my"][i, ..., s])
            l1 = _forced_abs(warped - target_x, ch["l1"][i, :, s].to(dt)).mean(1, keepdim=True)
            sv = _forced_ssim(warped, target_x, ch["pass"][i, :, s]).mean(1, keepdim=True)
            pes.append(alpha * sv + (1.0 - alpha) * l1)
        sel = ch["sel"][i].unsqueeze(1)                                   # (N,1,H,W): -1 = automask
        stack = torch.cat(pes, 1)
        warp_loss = torch.gather(stack, 1, sel.clamp(min=0))
        if auto_loss is not None:
            warp_loss = torch.where(sel < 0, auto_loss.to(dt), warp_loss)
        norm_disp = (disparity / (disparity.mean(dim=(2, 3), keepdim=True) + 1e-7))[:, 0]
        ddx = _forced_abs(norm_disp[:, :, :-1] - norm_disp[:, :, 1:], ch["smx"][i][:, :, :-1].to(dt))
        ddy = _forced_abs(norm_disp[:, :-1, :] - norm_disp[:, 1:, :], ch["smy"][i][:, :-1, :].to(dt))
        idx = (target_x[:, :, :, :-1] - target_x[:, :, :, 1:]).abs().mean(1)
        idy = (target_x[:, :, :-1, :] - target_x[:, :, 1:, :]).abs().mean(1)
        disp_loss = ((ddx * torch.exp(-idx)).mean() + (ddy * torch.exp(-idy)).mean()) * disparity_smoothness * scale
        loss = loss + warp_loss.mean() + disp_loss
    return loss / len(scales)


# ----------------------------------------------------------------------------
# Flux.Optimise.ADAM (Flux v0.12 optimise/optimisers.jl, `apply!(o::ADAM, x, Δ)`; third-party, restated from its
# published rule) and the triplet optimiser loop of src/simple_depth.jl:1-62
# ----------------------------------------------------------------------------


class FluxAdam:
    """mt = b1 mt + (1-b1) g;  vt = b2 vt + (1-b2) g^2;  x -= mt / (1-b1^t) / (sqrt(vt / (1-b2^t)) + eps) * eta"""

    def __init__(self, params, eta=1e-3, beta=(0.9, 0.999), eps=1e-8):
        self.params, self.eta, self.beta, self.eps = list(params), eta, beta, eps
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.bp = [beta[0], beta[1]]

    def step(self, grads):
        b1, b2 = self.beta
        with torch.no_grad():
            for p, g, m, v in zip(self.params, grads, self.m, self.v):
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).add_(g * g, alpha=1 - b2)
                p.sub_(m / (1 - self.bp[0]) / ((v / (1 - self.bp[1])).sqrt() + self.eps) * self.eta)
        self.bp = [self.bp[0] * b1, self.bp[1] * b2]


def slow_depth(x, K, invK, *, target_id=1, source_ids=(0, 2), min_depth=0.1, max_depth=100.0, iters=500, eta=3e-4):
    """src/simple_depth.jl:1-62 without the PNG dumps: x (1,L,C,H,W).  Returns (disp, rvecs, tvecs, loss history)."""
    dt = x.dtype
    N, L, C, H, W = x.shape
    disp = torch.full((N, 1, H, W), 0.5, dtype=dt, requires_grad=True)
    rvecs = [torch.tensor([[0.0, 0.0, 0.01]] * N, dtype=dt, requires_grad=True) for _ in source_ids]
    tvecs = [torch.zeros(N, 3, dtype=dt, requires_grad=True) for _ in source_ids]
    theta = [disp] + [t for pair in zip(rvecs, tvecs) for t in pair]
    opt = FluxAdam(theta, eta=eta)
    hist = []
    for _ in range(iters):
        loss = simple_depth_loss(x, disp, rvecs, tvecs, K, invK, target_id=target_id, source_ids=source_ids,
                                 min_depth=min_depth, max_depth=max_depth)
        grads = torch.autograd.grad(loss, theta)
        opt.step(grads)
        hist.append(float(loss.detach()))
    return disp.detach(), [r.detach() for r in rvecs], [t.detach() for t in tvecs], hist
