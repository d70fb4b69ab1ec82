"""Host-side mirror of the reference's operator interface for the view-synthesis loss path
(src/utils.jl, src/training.jl:1-19 of pxl-th/Monodepth2.jl): same names, argument meaning
and error behaviour, each backed by a hand-written CUDA kernel pair (forward + backward) in
libmd2_b200.so and differentiable through torch.autograd (standing in for Zygote +
ChainRulesCore rrules; the Julia binding with real rrules is julia/Monodepth2B200.jl).

Tensor conventions (row-major torch == column-major Julia memory):
  image (N,C,H,W) == Julia (W,H,C,N);  x (N,L,C,H,W) == (W,H,C,L,N);  disparity (N,1,H,W);
  points (N,P,3) == (3,P,N);  uv (N,P,2) == (2,P,N);  grid (N,H,W,2) == (2,W,H,N);
  K (3,3), R (N,3,3) [R[n,i,j], natural maths order], t (N,3) == (3,1,N), rvec (N,3) == (3,N).
CUDA float32 tensors only -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from ._lib import Context, require_cuda

_F32 = torch.float32


def _f32c(t):
    if t.dtype != _F32:
        raise TypeError(f"expected float32 CUDA tensor, got {t.dtype}")
    return t.contiguous()


def _cm(m):
    """(...,3,3) natural -> column-major (Julia) memory"""
    return m.transpose(-1, -2).contiguous()


def _ctx(t):
    require_cuda(t)
    return Context.get(t.device)


def _p(t):
    return None if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------------
# A1 disparity_to_depth (src/utils.jl:175-179)
# ---------------------------------------------------------------------------------------------
class _DispToDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, min_depth, max_depth):
        disp = _f32c(disp)
        out = torch.empty_like(disp)
        _ctx(disp).call("md2_disparity_to_depth_fwd", _p(disp), _p(out), disp.numel(), min_depth, max_depth)
        ctx.save_for_backward(disp)
        ctx.mm = (min_depth, max_depth)
        return out

    @staticmethod
    def backward(ctx, g):
        (disp,) = ctx.saved_tensors
        g = _f32c(g)
        gd = torch.empty_like(disp)
        _ctx(disp).call("md2_disparity_to_depth_bwd", _p(disp), _p(g), _p(gd), disp.numel(), *ctx.mm)
        return gd, None, None


def disparity_to_depth(disparity, min_depth, max_depth):
    return _DispToDepth.apply(disparity, float(min_depth), float(max_depth))


# ---------------------------------------------------------------------------------------------
# A2 Backproject (src/utils.jl:41-65)
# ---------------------------------------------------------------------------------------------
class _Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, invK_cm, W, H):
        depth = _f32c(depth)
        N = depth.numel() // (W * H)
        pts = torch.empty(N, W * H, 3, device=depth.device, dtype=_F32)
        _ctx(depth).call("md2_backproject_fwd", _p(depth), _p(invK_cm), _p(pts), W, H, N)
        ctx.save_for_backward(invK_cm)
        ctx.dims = (W, H, N, depth.shape)
        return pts

    @staticmethod
    def backward(ctx, g):
        (invK_cm,) = ctx.saved_tensors
        W, H, N, shape = ctx.dims
        g = _f32c(g)
        gd = torch.empty(shape, device=g.device, dtype=_F32)
        _ctx(g).call("md2_backproject_bwd", _p(g), _p(invK_cm), _p(gd), W, H, N)
        return gd, None, None, None


class Backproject:
    """Backproject(; width, height)(depth, invK): depth (N,P) [any shape with N*P elements,
    batch first] -> camera points (N,P,3)."""

    def __init__(self, width, height):
        self.width, self.height = int(width), int(height)

    def __call__(self, depth, invK):
        return _Backproject.apply(depth, _cm(_f32c(invK)), self.width, self.height)


# ---------------------------------------------------------------------------------------------
# A3 Project (src/utils.jl:67-99)
# ---------------------------------------------------------------------------------------------
class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, K_cm, R, t, W, H):
        points, t = _f32c(points), _f32c(t)
        R_cm = _cm(_f32c(R))
        N = points.shape[0]
        uv = torch.empty(N, W * H, 2, device=points.device, dtype=_F32)
        _ctx(points).call("md2_project_fwd", _p(points), _p(K_cm), _p(R_cm), _p(t), _p(uv), W, H, N)
        ctx.save_for_backward(points, K_cm, R_cm, t)
        ctx.dims = (W, H, N)
        return uv

    @staticmethod
    def backward(ctx, g):
        points, K_cm, R_cm, t = ctx.saved_tensors
        W, H, N = ctx.dims
        g = _f32c(g)
        gp = torch.empty_like(points)
        gR_cm = torch.empty_like(R_cm)
        gt = torch.empty_like(t)
        _ctx(g).call("md2_project_bwd", _p(points), _p(K_cm), _p(R_cm), _p(t), _p(g), _p(gp), _p(gR_cm), _p(gt),
                     W, H, N)
        return gp, None, gR_cm.transpose(1, 2), gt, None, None


class Project:
    """Project(; width, height)(points, K, R, t) -> (N,P,2) normalised to (-1,1).
    K may be (3,3) or (1,3,3) (test/runtests.jl:104 passes (3,3,1))."""

    def __init__(self, width, height):
        self.width, self.height = int(width), int(height)

    def __call__(self, points, K, R, t):
        K = K.reshape(3, 3)
        return _Project.apply(points, _cm(_f32c(K)), R, t.reshape(-1, 3), self.width, self.height)


# ---------------------------------------------------------------------------------------------
# A4-A6 so3_exp_map / hat / composeT (src/utils.jl:101-141, 181-188)
# ---------------------------------------------------------------------------------------------
class _So3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rvec):
        rvec = _f32c(rvec)
        N = rvec.shape[0]
        R_cm = torch.empty(N, 3, 3, device=rvec.device, dtype=_F32)
        _ctx(rvec).call("md2_so3_exp_map_fwd", _p(rvec), _p(R_cm), N)
        ctx.save_for_backward(rvec)
        return R_cm.transpose(1, 2)

    @staticmethod
    def backward(ctx, g):
        (rvec,) = ctx.saved_tensors
        g_cm = _cm(_f32c(g))
        out = torch.empty_like(rvec)
        _ctx(rvec).call("md2_so3_exp_map_bwd", _p(rvec), _p(g_cm), _p(out), rvec.shape[0])
        return out


def so3_exp_map(rvec):
    return _So3.apply(rvec)


class _Hat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rvec):
        rvec = _f32c(rvec)
        N = rvec.shape[0]
        S_cm = torch.empty(N, 3, 3, device=rvec.device, dtype=_F32)
        _ctx(rvec).call("md2_hat_fwd", _p(rvec), _p(S_cm), N)
        return S_cm.transpose(1, 2)

    @staticmethod
    def backward(ctx, g):
        g_cm = _cm(_f32c(g))
        N = g_cm.shape[0]
        out = torch.empty(N, 3, device=g.device, dtype=_F32)
        _ctx(g_cm).call("md2_hat_bwd", _p(g_cm), _p(out), N)
        return out


def hat(rvec):
    return _Hat.apply(rvec)


class _ComposeT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rvec, tvec, invert):
        rvec, tvec = _f32c(rvec), _f32c(tvec)
        N = rvec.shape[0]
        R_cm = torch.empty(N, 3, 3, device=rvec.device, dtype=_F32)
        t = torch.empty(N, 3, device=rvec.device, dtype=_F32)
        _ctx(rvec).call("md2_compose_T_fwd", _p(rvec), _p(tvec), int(invert), _p(R_cm), _p(t), N)
        ctx.save_for_backward(rvec, tvec)
        ctx.invert = int(invert)
        return R_cm.transpose(1, 2), t

    @staticmethod
    def backward(ctx, gR, gt):
        rvec, tvec = ctx.saved_tensors
        gR_cm = _cm(_f32c(gR)) if gR is not None else None
        gt = _f32c(gt) if gt is not None else None
        grv, gtv = torch.empty_like(rvec), torch.empty_like(tvec)
        _ctx(rvec).call("md2_compose_T_bwd", _p(rvec), _p(tvec), ctx.invert, _p(gR_cm), _p(gt), _p(grv), _p(gtv),
                        rvec.shape[0])
        return grv, gtv, None


def composeT(rvec, t, invert):
    R, tu = _ComposeT.apply(rvec, t.reshape(-1, 3), bool(invert))
    return R, tu


# ---------------------------------------------------------------------------------------------
# A16 NNlib.grid_sample, A17 NNlib.upsample_bilinear
# ---------------------------------------------------------------------------------------------
_PAD = {"zeros": 0, "border": 1}


class _GridSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, grid, mode):
        inp, grid = _f32c(inp), _f32c(grid)
        N, Cc, H, W = inp.shape
        Ho, Wo = grid.shape[1], grid.shape[2]
        out = torch.empty(N, Cc, Ho, Wo, device=inp.device, dtype=_F32)
        _ctx(inp).call("md2_grid_sample_fwd", _p(inp), _p(grid), _p(out), W, H, Cc, N, Wo, Ho, mode)
        ctx.save_for_backward(inp, grid)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, g):
        inp, grid = ctx.saved_tensors
        g = _f32c(g)
        N, Cc, H, W = inp.shape
        Ho, Wo = grid.shape[1], grid.shape[2]
        gi = torch.zeros_like(inp) if ctx.needs_input_grad[0] else None
        gg = torch.empty_like(grid) if ctx.needs_input_grad[1] else None
        _ctx(inp).call("md2_grid_sample_bwd", _p(inp), _p(grid), _p(g), _p(gi), _p(gg), W, H, Cc, N, Wo, Ho, ctx.mode)
        return gi, gg, None


def grid_sample(input, grid, padding_mode="zeros"):
    if padding_mode not in _PAD:
        raise ValueError("padding_mode must be 'zeros' or 'border'")
    return _GridSample.apply(input, grid, _PAD[padding_mode])


class _Upsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, H):
        x = _f32c(x)
        N, Cc, h, w = x.shape
        out = torch.empty(N, Cc, H, W, device=x.device, dtype=_F32)
        _ctx(x).call("md2_upsample_bilinear_fwd", _p(x), _p(out), w, h, W, H, N * Cc)
        ctx.dims = (N, Cc, h, w, W, H)
        return out

    @staticmethod
    def backward(ctx, g):
        N, Cc, h, w, W, H = ctx.dims
        g = _f32c(g)
        gi = torch.empty(N, Cc, h, w, device=g.device, dtype=_F32)
        _ctx(g).call("md2_upsample_bilinear_bwd", _p(g), _p(gi), w, h, W, H, N * Cc)
        return gi, None, None


def upsample_bilinear(x, size):
    """size = (W, H) like NNlib.upsample_bilinear(x; size=(width, height))"""
    return _Upsample.apply(x, int(size[0]), int(size[1]))


# ---------------------------------------------------------------------------------------------
# A7 SSIM (src/utils.jl:13-39)
# ---------------------------------------------------------------------------------------------
class _Ssim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = _f32c(x), _f32c(y)
        if x.shape != y.shape:
            raise ValueError("SSIM: x and y must have the same shape")
        N, Cc, H, W = x.shape
        out = torch.empty_like(x)
        _ctx(x).call("md2_ssim_fwd", _p(x), _p(y), _p(out), W, H, Cc, N)
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        g = _f32c(g)
        N, Cc, H, W = x.shape
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        _ctx(x).call("md2_ssim_bwd", _p(x), _p(y), _p(g), _p(gx), _p(gy), W, H, Cc, N)
        return gx, gy


class SSIM:
    """SSIM()(x, y): dissimilarity clamp((1 - SSIM)/2, 0, 1), 3x3 mean pool over reflect-padded
    input, c1 = 0.01^2, c2 = 0.03^2."""
    c1, c2 = 0.01 ** 2, 0.03 ** 2

    def __call__(self, x, y):
        return _Ssim.apply(x, y)


# ---------------------------------------------------------------------------------------------
# A10-A13 photometric_loss / prediction_loss / automasking_loss / _apply_mask
# (src/training.jl:1-19)
# ---------------------------------------------------------------------------------------------
def _ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def _photomin_fwd(preds, strides, target, target_stride, mask, alpha, shape, want_argmin):
    N, Cc, H, W = shape
    out = torch.empty(N, 1, H, W, device=target.device, dtype=_F32)
    argmin = torch.empty(N, 1, H, W, device=target.device, dtype=torch.int32) if want_argmin else None
    S = len(preds)
    _ctx(target).call("md2_photometric_min_fwd", S, _ptr_array([p.data_ptr() for p in preds]),
                      (C.c_int64 * S)(*strides), _p(target), target_stride, _p(mask), float(alpha), _p(out),
                      _p(argmin), W, H, Cc, N)
    return out, argmin


class _PhotoMin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, target, *preds):
        target = _f32c(target)
        preds = [_f32c(p) for p in preds]
        for p in preds:
            if p.shape != target.shape:
                raise ValueError("prediction and target shapes differ")
        N, Cc, H, W = target.shape
        chw = Cc * H * W
        out, argmin = _photomin_fwd(preds, [chw] * len(preds), target, chw, None, alpha, target.shape, True)
        ctx.save_for_backward(target, argmin, *preds)
        ctx.alpha = alpha
        return out

    @staticmethod
    def backward(ctx, g):
        target, argmin, *preds = ctx.saved_tensors
        g = _f32c(g)
        N, Cc, H, W = target.shape
        S, chw = len(preds), Cc * H * W
        gt = torch.empty_like(target) if ctx.needs_input_grad[1] else None
        gp = [torch.empty_like(p) if ctx.needs_input_grad[2 + i] else None for i, p in enumerate(preds)]
        _ctx(target).call("md2_photometric_min_bwd", S, _ptr_array([p.data_ptr() for p in preds]),
                          (C.c_int64 * S)(*([chw] * S)), _p(target), chw, None, float(ctx.alpha), _p(g), _p(argmin),
                          _ptr_array([_p(t) for t in gp]), _p(gt), None, W, H, Cc, N)
        return (None, gt, *gp)


def photometric_loss(ssim, predicted, target, alpha=0.85):
    """alpha * mean_c SSIM(predicted, target) + (1 - alpha) * mean_c |target - predicted|"""
    return _PhotoMin.apply(float(alpha), target, predicted)


def prediction_loss(ssim, predictions, target):
    """per-pixel minimum over the predictions of photometric_loss (first index wins ties)"""
    return _PhotoMin.apply(0.85, target, *predictions)


def automasking_loss(ssim, inputs, target, source_ids):
    """min over the UN-warped source frames; inputs (N,L,C,H,W), source_ids 0-based.  A constant
    in the reference (computed outside `gradient`, src/Monodepth.jl:159-164): forward only."""
    inputs, target = _f32c(inputs.detach()), target.detach()
    N, Lf, Cc, H, W = inputs.shape
    if target.dtype != _F32 or target.stride()[1:] != (H * W, W, 1):
        target = _f32c(target)
    preds = [inputs[:, i] for i in source_ids]
    out, _ = _photomin_fwd(preds, [inputs.stride(0)] * len(preds), target, target.stride(0), None, 0.85,
                           (N, Cc, H, W), False)
    return out


def _apply_mask(mask, warp_loss):
    """minimum(cat(mask, warp_loss; dims=3); dims=3): the mask is first, so it wins ties and
    takes the gradient there.  Element-wise glue, left to the host framework."""
    return torch.where(mask <= warp_loss, mask, warp_loss)


# ---------------------------------------------------------------------------------------------
# A8 smooth_loss (src/utils.jl:143-173)
# ---------------------------------------------------------------------------------------------
class _Smooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, image, normalize):
        disp, image = _f32c(disp), _f32c(image)
        N, Cc, H, W = image.shape
        if disp.numel() != N * H * W:
            raise ValueError("smooth_loss: disparity must be (N,H,W) matching the image")
        out = torch.empty((), device=disp.device, dtype=_F32)
        _ctx(disp).call("md2_smooth_loss_fwd", _p(disp), _p(image), Cc * H * W, _p(out), int(normalize), W, H, Cc, N)
        ctx.save_for_backward(disp, image)
        ctx.normalize = int(normalize)
        return out

    @staticmethod
    def backward(ctx, g):
        disp, image = ctx.saved_tensors
        N, Cc, H, W = image.shape
        gd = torch.empty_like(disp)
        gi = torch.empty_like(image) if ctx.needs_input_grad[1] else None
        _ctx(disp).call("md2_smooth_loss_bwd", _p(disp), _p(image), Cc * H * W, 1.0, _p(gd), _p(gi), ctx.normalize,
                        W, H, Cc, N)
        return gd * g, (gi * g if gi is not None else None), None


def smooth_loss(disparity, image, normalize=False):
    """disparity (N,H,W) [or (N,1,H,W)], image (N,C,H,W) -> scalar.  normalize=True folds in the
    d / (mean d + 1e-7) of src/training.jl:64-65."""
    return _Smooth.apply(disparity, image, bool(normalize))
