"""CPU ORACLE no. 2, Python face (test infrastructure, NOT product code): ctypes loader of oracle/libc_oracle.so, the
plain-C float64 restatement of the path with a hand-derived reverse pass (oracle/c_oracle.c).  Same argument conventions
as oracle/torch_oracle.py, so a test can hand the same tensors to both.  Only tests/, __graft_entry__ and bench.py's
CPU arms may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libc_oracle.so")
_lib = None
_D = C.POINTER(C.c_double)


def build():
    src = os.path.join(HERE, "c_oracle.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", HERE, "libc_oracle.so"], check=True)
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.co_smooth_loss.restype = C.c_double
        _lib.co_view_synthesis_loss.restype = C.c_int
    return _lib


def _f64(t):
    return t.detach().to(torch.float64).contiguous()


def _p(t):
    return C.cast(t.data_ptr(), _D) if t is not None else None


def disparity_to_depth(d, min_depth, max_depth):
    d = _f64(d)
    out = torch.empty_like(d)
    lib().co_disparity_to_depth(_p(d), C.c_long(d.numel()), C.c_double(min_depth), C.c_double(max_depth), _p(out))
    return out


def so3_exp_map(rvec):
    r = _f64(rvec)
    R = torch.empty(r.shape[0], 3, 3, dtype=torch.float64)
    lib().co_so3_exp_map(_p(r), r.shape[0], _p(R))
    return R


def so3_exp_map_bwd(rvec, Rb):
    r, Rb = _f64(rvec), _f64(Rb)
    rb = torch.empty_like(r)
    lib().co_so3_exp_map_bwd(_p(r), r.shape[0], _p(Rb), _p(rb))
    return rb


def composeT(rvec, t, invert):
    r, t = _f64(rvec), _f64(t)
    R = torch.empty(r.shape[0], 3, 3, dtype=torch.float64)
    to = torch.empty_like(t)
    lib().co_composeT(_p(r), _p(t), r.shape[0], int(bool(invert)), _p(R), _p(to))
    return R, to


def ssim(x, y):
    """x, y (N,C,H,W) -> (N,C,H,W)"""
    x, y = _f64(x), _f64(y)
    N, Cc, H, W = x.shape
    out = torch.empty_like(x)
    lib().co_ssim(_p(x), _p(y), N * Cc, H, W, _p(out))
    return out


def smooth_loss(disparity, image, grad=False):
    """disparity (N,H,W), image (N,C,H,W) -> loss [, d loss / d disparity]"""
    d, im = _f64(disparity), _f64(image)
    N, Cc, H, W = im.shape
    g = torch.zeros_like(d) if grad else None
    v = lib().co_smooth_loss(_p(d), _p(im), N, Cc, H, W, _p(g))
    return (v, g) if grad else v


def upsample_bilinear(x, size_wh):
    x = _f64(x)
    N, Cc, h, w = x.shape
    W, H = size_wh
    out = torch.empty(N, Cc, H, W, dtype=torch.float64)
    lib().co_upsample_bilinear(_p(x), N * Cc, h, w, H, W, _p(out))
    return out


def grid_sample_border(inp, grid):
    """inp (N,C,H,W), grid (N,H,W,2) -> (N,C,H,W)"""
    inp, grid = _f64(inp), _f64(grid)
    N, Cc, H, W = inp.shape
    out = torch.empty_like(inp)
    for n in range(N):
        lib().co_grid_sample_border(_p(inp[n]), _p(grid[n]), Cc, H, W, _p(out[n]))
    return out


def view_synthesis_loss(x, disparities, rvecs, tvecs, K, invK, *, target_id=1, source_ids=(0, 2),
                        scales=(0.125, 0.25, 0.5, 1.0), min_depth=0.1, max_depth=100.0, disparity_smoothness=1e-3,
                        auto_loss=None, normalize_disp=True, alpha=0.85, grad=True, viz=False, choices=None,
                        export_choices=False):
    """Value and hand-derived gradients of the train_loss tail (src/training.jl:42-78).  Returns a dict with loss, gdisp,
    grvec, gtvec, gx (source frames only: the target frame is data, as in the fused CUDA call) [, viz_warped, viz_loss].
    `choices` (int32 (L,N,H,W,1+S), md2_vsl_desc.debug_choices): evaluate with the discrete decisions of the implementation
    under test forced, like torch_oracle.view_synthesis_loss_forced.  export_choices: also return this evaluation's own
    decisions in that layout (out["choices"])."""
    x = _f64(x)
    N, L, Cc, H, W = x.shape
    S, n = len(source_ids), len(disparities)
    ds = [_f64(d) for d in disparities]
    rv = torch.stack([_f64(r) for r in rvecs]).contiguous()
    tv = torch.stack([_f64(t) for t in tvecs]).contiguous()
    Kd, iKd = _f64(K).reshape(3, 3).contiguous(), _f64(invK).reshape(3, 3).contiguous()
    al = _f64(auto_loss).reshape(N, H, W) if auto_loss is not None else None
    gd = [torch.zeros_like(d) for d in ds]
    grv, gtv, gx = torch.zeros_like(rv), torch.zeros_like(tv), torch.zeros_like(x)
    vw = torch.zeros(S, N, Cc, H, W, dtype=torch.float64) if viz else None
    vl = torch.zeros(N, H, W, dtype=torch.float64) if viz else None
    loss = C.c_double(0.0)
    chp = None
    if choices is not None:
        choices = choices.detach().cpu().to(torch.int32).contiguous()
        assert tuple(choices.shape) == (n, N, H, W, 1 + S), choices.shape
        chp = C.cast(choices.data_ptr(), C.POINTER(C.c_int))
    cho = torch.zeros(n, N, H, W, 1 + S, dtype=torch.int32) if export_choices else None
    chop = C.cast(cho.data_ptr(), C.POINTER(C.c_int)) if export_choices else None
    arr = lambda ts: (_D * n)(*[_p(t) for t in ts])
    ints = lambda v: (C.c_int * len(v))(*v)
    sc = (C.c_double * n)(*[float(s) for s in scales[:n]])
    rc = lib().co_view_synthesis_loss(
        _p(x), N, L, Cc, H, W, n, arr(ds), ints([d.shape[-2] for d in ds]), ints([d.shape[-1] for d in ds]),
        _p(rv), _p(tv), _p(Kd), _p(iKd), int(target_id), S, ints(list(source_ids)), sc,
        C.c_double(min_depth), C.c_double(max_depth), C.c_double(disparity_smoothness), _p(al),
        int(bool(normalize_disp)), C.c_double(alpha), C.byref(loss), _p(gx) if grad else None,
        arr(gd) if grad else None, _p(grv) if grad else None, _p(gtv) if grad else None, _p(vw), _p(vl), chp, chop)
    if rc != 0:
        raise ValueError(f"co_view_synthesis_loss: status {rc}")
    out = {"loss": loss.value}
    if grad:
        out.update(gdisp=gd, grvec=list(grv), gtvec=list(gtv), gx=gx)
    if viz:
        out.update(viz_warped=list(vw), viz_loss=vl.unsqueeze(1))
    if export_choices:
        out["choices"] = cho
    return out


def simple_depth_loss(x, disp, rvecs, tvecs, K, invK, **kw):
    """src/simple_depth.jl:25-41: one scale, un-normalised smoothness of weight 1"""
    return view_synthesis_loss(x, [disp], rvecs, tvecs, K, invK, scales=(1.0,), disparity_smoothness=1.0,
                               normalize_disp=False, **kw)
