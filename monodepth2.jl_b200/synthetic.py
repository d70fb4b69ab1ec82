"""Seeded synthetic KITTI-shaped inputs for the view-synthesis loss path: what
the reference's datasets (src/kitty.jl, src/dtk.jl) and networks hand to `train_loss`, with no
files involved.  Host-side data only -- nothing here computes the loss.

The tests check that these generators produce exactly the data of the test-side generator, so
the CUDA arm, the CPU arm and the parity tests all see identical inputs.
"""
from __future__ import annotations

import torch


def make_K(W, H, f=None, dtype=torch.float32):
    """KITTI-shaped intrinsics in the reference's 1-based pixel frame: f = 0.58 W, cx = W/2,
    cy = H/2 (src/kitty.jl:27-28); returns (K, K^-1) as row-major 3x3"""
    f = 0.58 * W if f is None else f
    K = torch.tensor([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    return K.to(dtype), torch.linalg.inv(K).to(dtype)


def scale_sizes(W, H, scales=(0.125, 0.25, 0.5, 1.0)):
    """(w, h) of the decoder's disparity maps (src/depth_decoder.jl: scale_levels 2:5)"""
    return [(max(2, int(round(W * s))), max(2, int(round(H * s)))) for s in scales]


def synthetic_batch(N, C, H, W, scales=(0.125, 0.25, 0.5, 1.0), seed=42, full_res_disp=False,
                    pose_sigma=0.01, dtype=torch.float32):
    """x (N,3,C,H,W) in [0,1]: smooth textured target, sources = target shifted by up to 3 px,
    + 5 % noise; disparities: low-passed sigmoid noise at the native size of every scale;
    poses: rvec, tvec ~ pose_sigma N(0,1) with |rvec| kept away from 0 (README.md:47-51)."""
    g = torch.Generator().manual_seed(seed)
    f64 = torch.float64
    rows = torch.arange(H, dtype=f64).view(H, 1).expand(H, W)
    cols = torch.arange(W, dtype=f64).view(1, W).expand(H, W)
    waves = 6
    x = torch.zeros(N, 3, C, H, W, dtype=f64)
    for n in range(N):
        offs = []
        for frame in range(3):
            o = (torch.rand(2, generator=g, dtype=f64) - 0.5) * 6.0
            offs.append(torch.zeros(2, dtype=f64) if frame == 1 else o)
        for c in range(C):
            kx = (torch.rand(waves, generator=g, dtype=f64) - 0.5) * 0.5
            ky = (torch.rand(waves, generator=g, dtype=f64) - 0.5) * 0.5
            phase = torch.rand(waves, generator=g, dtype=f64) * 6.283
            amp = torch.rand(waves, generator=g, dtype=f64) / waves
            for frame in range(3):
                field = torch.zeros(H, W, dtype=f64)
                for j in range(waves):
                    field = field + amp[j] * torch.sin(kx[j] * (cols + offs[frame][0]) + ky[j] * (rows + offs[frame][1]) + phase[j])
                x[n, frame, c] = 0.5 + 0.45 * field
    x = (x + 0.05 * (torch.rand(x.shape, generator=g, dtype=f64) - 0.5)).clamp(0, 1)
    disps = []
    for s in scales:
        h, w = (H, W) if full_res_disp else (max(2, int(round(H * s))), max(2, int(round(W * s))))
        d = torch.randn(N, 1, h, w, generator=g, dtype=f64)
        d = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(d, (2, 2, 2, 2), mode="replicate"), 5, 1)
        disps.append(torch.sigmoid(2.0 * d).to(dtype))
    rvecs, tvecs = [], []
    for _ in range(2):
        r = pose_sigma * torch.randn(N, 3, generator=g, dtype=f64)
        small = r.norm(dim=1, keepdim=True) < 1e-3
        r = torch.where(small, r + 2e-3, r)
        rvecs.append(r.to(dtype))
        tvecs.append((pose_sigma * torch.randn(N, 3, generator=g, dtype=f64)).to(dtype))
    return x.to(dtype), disps, rvecs, tvecs


def algorithmic_bytes(W, H, N, C, S, L, m=0, g=1):
    """compulsory fp32 traffic of one fused forward+backward step:
    fwd 4(1+C+SC+m) + bwd 4(1+C+SC+m) + 4(1+gSC) bytes per unit (pixel x scale x image);
    returns (bytes per step, bytes per unit)"""
    per_unit = 4 * (1 + C + S * C + m) * 2 + 4 * (1 + g * S * C)
    return per_unit * W * H * N * L, per_unit


__all__ = ["make_K", "scale_sizes", "synthetic_batch", "algorithmic_bytes"]
