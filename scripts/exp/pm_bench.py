#!/usr/bin/env python
"""time of the automask pre-pass (photometric_min forward) at BASELINE's automask configurations, cold L2 (256 MB written
between repetitions) and warm; usage: python scripts/exp/pm_bench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import monodepth2_jl_b200 as M  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(64 << 20, device=dev)
for (W, H, N, C) in [(640, 192, 12, 3), (416, 128, 64, 3), (416, 128, 8, 1), (1024, 320, 4, 3)]:
    x = torch.rand(N, 3, C, H, W, device=dev)
    ssim = M.SSIM()
    fn = lambda: M.automasking_loss(ssim, x, x[:, 1], (0, 2))
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        fn()
    b.record()
    torch.cuda.synchronize()
    warm = a.elapsed_time(b) / 50
    cold = 0.0
    for _ in range(20):
        flush.fill_(1.0)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        cold += a.elapsed_time(b) / 20
    byts = (3 * C + 1) * N * H * W * 4
    print(f"{W}x{H}x{N} C={C}: warm {warm * 1e3:.1f} us, cold {cold * 1e3:.1f} us = {byts / cold / 1e6:.0f} GB/s", flush=True)
