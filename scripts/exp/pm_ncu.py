#!/usr/bin/env python
"""launch the automask pre-pass a few times on one shape (for `ncu --metrics gpu__time_duration.sum`: per-launch kernel
times without the Python glue); usage: python scripts/exp/pm_ncu.py W H N C"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import monodepth2_jl_b200 as M  # noqa: E402

W, H, N, C = [int(v) for v in sys.argv[1:5]]
dev = torch.device("cuda", 0)
x = torch.rand(N, 3, C, H, W, device=dev)
for _ in range(12):
    M.automasking_loss(M.SSIM(), x, x[:, 1], (0, 2))
torch.cuda.synchronize()
