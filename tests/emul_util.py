"""Test helper: drive tests/emul/libmd2_emul.so (host emulation of the fused CUDA tile phases)."""
import ctypes as C
import os
import subprocess

import torch

from monodepth2_jl_b200 import _lib as L

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL_DIR = os.path.join(HERE, "emul")
EMUL_SO = os.path.join(EMUL_DIR, "libmd2_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "monodepth2.jl_b200", "csrc")


def build_emul():
    srcs = [os.path.join(EMUL_DIR, "md2_emul.cpp"), os.path.join(EMUL_DIR, "warp_emu.h"), os.path.join(CSRC, "md2_fused.cuh"),
            os.path.join(CSRC, "md2_march.cuh"), os.path.join(CSRC, "md2_march2.cuh"), os.path.join(CSRC, "md2_math.cuh")]
    if os.path.exists(EMUL_SO) and all(os.path.getmtime(s) <= os.path.getmtime(EMUL_SO) for s in srcs):
        return EMUL_SO
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", EMUL_SO,
                    os.path.join(EMUL_DIR, "md2_emul.cpp")], check=True)
    return EMUL_SO


def emul_vsl(x, disps, rvecs, tvecs, K, invK, *, mode=2, gloss=1.0, target_id=1, source_ids=(0, 2),
             scales=(0.125, 0.25, 0.5, 1.0), min_depth=0.1, max_depth=100.0, disparity_smoothness=1e-3,
             automask=None, normalize=True, smooth_weight=None, loss_scale=None, grad_source=True, saved=None,
             viz=False, variant="march2", R=32, debug_choices=False):
    """CPU float32 tensors in, dict of outputs out (same maths as the CUDA fused path)."""
    lib = C.CDLL(build_emul())
    lib.md2_emul_vsl.argtypes = [C.POINTER(L.VslDesc), C.c_int, C.c_float]
    lib.md2_emul_march.argtypes = [C.POINTER(L.VslDesc), C.c_int, C.c_float, C.c_int]
    lib.md2_emul_march2.argtypes = [C.POINTER(L.VslDesc), C.c_int, C.c_float, C.c_int]
    N, F, Cc, H, W = x.shape
    S, Ls = len(source_ids), len(disps)
    x = x.contiguous()
    disps = [d.contiguous() for d in disps]
    rv = [r.contiguous() for r in rvecs]
    tv = [t.contiguous() for t in tvecs]
    K_cm, invK_cm = K.t().contiguous(), invK.t().contiguous()
    out = {
        "loss": torch.zeros(1),
        "gdisp": [torch.zeros_like(d) for d in disps],
        "grvec": [torch.zeros_like(r) for r in rv],
        "gtvec": [torch.zeros_like(t) for t in tv],
        "gx": torch.zeros_like(x) if grad_source else None,
        "saved": torch.zeros(Ls, N, 4) if saved is None else saved,
    }
    if viz:
        out["viz_warped"] = [torch.zeros(N, Cc, H, W) for _ in range(S)]
        out["viz_loss"] = torch.zeros(N, 1, H, W)
    if debug_choices:
        out["choices"] = torch.zeros(Ls, N, H, W, 1 + S, dtype=torch.int32)
    sw = smooth_weight if smooth_weight is not None else [disparity_smoothness * s for s in scales[:Ls]]
    ls = loss_scale if loss_scale is not None else 1.0 / Ls
    desc = L.make_vsl_desc(
        target=x[:, target_id], target_stride=x.stride(0),
        sources=[x[:, i] for i in source_ids], source_strides=[x.stride(0)] * S,
        disparities=disps, K_cm=K_cm, invK_cm=invK_cm, rot=rv, trans=tv, pose_mode=1,
        invert=[i < target_id for i in source_ids], automask=automask, min_depth=min_depth, max_depth=max_depth,
        smooth_weight=sw, loss_scale=ls, normalize_disparity=normalize, loss=out["loss"],
        grad_disparity=out["gdisp"], grad_rot=out["grvec"], grad_trans=out["gtvec"],
        grad_source=[out["gx"][:, i] for i in source_ids] if grad_source else None,
        viz_warped=out.get("viz_warped"), viz_loss=out.get("viz_loss"), saved=out["saved"], debug_choices=out.get("choices"),
        shape=(N, Cc, H, W))
    if variant == "march2":
        rc = lib.md2_emul_march2(C.byref(desc), mode, gloss, R)
    elif variant == "march":
        rc = lib.md2_emul_march(C.byref(desc), mode, gloss, R)
    else:
        rc = lib.md2_emul_vsl(C.byref(desc), mode, gloss)
    assert rc == 0
    return out
