"""Generates the committed golden fixtures (run in the build container, where /root/reference
exists; the GPU box only reads the .npz files).

  python tests/golden/make_golden.py

1. simple_depth_c1.npz -- config 1 (BASELINE.json configs[0]): the reference's own Depth10k
   triplet res/image.png (1248x128 RGB = 3 x 416x128, src/dtk.jl:16-47), the slow_depth start
   point (disp = 0.5, rvec = [0,0,0.01], tvec = 0; src/simple_depth.jl:8-14) and the oracle's
   float64 loss / gradients for the objective of src/simple_depth.jl:25-41.
2. vsl_small.npz -- a seeded synthetic 4-scale train_loss case with automasking: inputs and the
   oracle's float64 loss / gradients.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import torch_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
F64 = torch.float64


def simple_depth_c1():
    from PIL import Image
    im = np.asarray(Image.open("/root/reference/res/image.png").convert("RGB"))  # (128,1248,3)
    W, H = 416, 128
    frames = np.stack([im[:, W * j:W * (j + 1)] for j in range(3)], 0)  # (3,H,W,3) uint8
    x = torch.from_numpy(frames).permute(0, 3, 1, 2).to(F64).div(255.0).unsqueeze(0)  # (1,3,C,H,W)
    focal = 2648.0 / 4.63461538462
    K, invK = O.make_K(W, H, f=focal, dtype=F64)
    disp = torch.full((1, 1, H, W), 0.5, dtype=F64, requires_grad=True)
    rv = [torch.tensor([[0.0, 0.0, 0.01]], dtype=F64, requires_grad=True) for _ in range(2)]
    tv = [torch.zeros(1, 3, dtype=F64, requires_grad=True) for _ in range(2)]
    loss = O.simple_depth_loss(x, disp, rv, tv, K, invK)
    loss.backward()
    # a second point away from the degenerate start (t = 0 makes d loss / d disp vanish):
    # smooth disparity bump and small translations, as after some optimiser steps
    yy, xx = torch.meshgrid(torch.arange(H, dtype=F64), torch.arange(W, dtype=F64), indexing="ij")
    disp2 = (0.5 + 0.2 * torch.sin(xx / 37.0) * torch.cos(yy / 23.0)).reshape(1, 1, H, W).requires_grad_(True)
    rv2 = [torch.tensor([[0.002, -0.004, 0.01]], dtype=F64, requires_grad=True),
           torch.tensor([[-0.003, 0.005, -0.008]], dtype=F64, requires_grad=True)]
    tv2 = [torch.tensor([[0.02, 0.003, -0.05]], dtype=F64, requires_grad=True),
           torch.tensor([[-0.015, -0.002, 0.06]], dtype=F64, requires_grad=True)]
    loss2 = O.simple_depth_loss(x, disp2, rv2, tv2, K, invK)
    loss2.backward()
    np.savez_compressed(
        os.path.join(HERE, "simple_depth_c1.npz"), frames=frames, focal=focal, loss=loss.item(),
        gdisp=disp.grad.numpy().astype(np.float32),
        grvec=np.stack([r.grad.numpy() for r in rv]), gtvec=np.stack([t.grad.numpy() for t in tv]),
        rvec2=np.stack([r.detach().numpy() for r in rv2]), tvec2=np.stack([t.detach().numpy() for t in tv2]),
        loss2=loss2.item(), gdisp2=disp2.grad.numpy().astype(np.float32),
        grvec2=np.stack([r.grad.numpy() for r in rv2]), gtvec2=np.stack([t.grad.numpy() for t in tv2]))
    print("simple_depth_c1: loss", loss.item(), "loss2", loss2.item())


def vsl_small():
    N, C, H, W = 2, 3, 48, 96
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=7)
    K, invK = O.make_K(W, H)
    xd = x.double()
    dd = [d.double().requires_grad_(True) for d in disps]
    rd = [r.double().requires_grad_(True) for r in rv]
    td = [t.double().requires_grad_(True) for t in tv]
    auto = O.automasking_loss(O.SSIM(), xd, xd[:, 1], (0, 2))
    loss = O.view_synthesis_loss(xd, dd, rd, td, K.double(), invK.double(), automasking=True, auto_loss=auto)
    loss.backward()
    out = dict(x=x.numpy(), K=K.numpy(), invK=invK.numpy(), loss=loss.item(), auto=auto.numpy().astype(np.float32))
    for i in range(4):
        out[f"disp{i}"] = disps[i].numpy()
        out[f"gdisp{i}"] = dd[i].grad.numpy()
    for s in range(2):
        out[f"rvec{s}"], out[f"tvec{s}"] = rv[s].numpy(), tv[s].numpy()
        out[f"grvec{s}"], out[f"gtvec{s}"] = rd[s].grad.numpy(), td[s].grad.numpy()
    np.savez_compressed(os.path.join(HERE, "vsl_small.npz"), **out)
    print("vsl_small: loss", loss.item())


if __name__ == "__main__":
    simple_depth_c1()
    vsl_small()
