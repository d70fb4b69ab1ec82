#!/usr/bin/env python
"""BASELINE.json configs[4]: micro-benchmark table of the stand-alone operators and the fused call, forward and
forward+backward, warm (same buffers back to back: L2-resident) and cold (256 MB written between repetitions), at
416x128 .. 1024x320 x batch 1 .. 64.  One JSON line per (op, shape); GB/s = compulsory bytes of the op (inputs read
once + outputs written once, fp32) over the cold time.  Usage: python scripts/microbench_ops.py [--quick]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import monodepth2_jl_b200 as M  # noqa: E402
from monodepth2_jl_b200 import synthetic as SY  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(64 << 20, device=dev)          # 256 MB > 2 x L2


def timeit(fn, reps=20):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    warm = a.elapsed_time(b) / reps
    for e0, e1 in ev:
        flush.fill_(1.0)
        e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    cold = sum(e0.elapsed_time(e1) for e0, e1 in ev) / reps
    return warm, cold


def main():
    quick = "--quick" in sys.argv
    shapes = [(416, 128, 1), (416, 128, 8), (416, 128, 64), (640, 192, 12), (1024, 320, 4)] if not quick else [(416, 128, 8)]
    peak = 6449.4
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    for (W, H, N) in shapes:
        C = 3
        x, disps, rv, tv = SY.synthetic_batch(min(N, 2), C, H, W, seed=3)
        rep = (N + 1) // 2
        tile = lambda t: t.repeat(rep, *([1] * (t.dim() - 1)))[:N].contiguous().to(dev)
        x, disps, rv, tv = tile(x), [tile(d) for d in disps], [tile(r) for r in rv], [tile(t) for t in tv]
        K, invK = [t.to(dev) for t in SY.make_K(W, H)]
        px = N * C * H * W * 4            # bytes of one (N,C,H,W) image
        p1 = N * H * W * 4
        a, b, c2 = x[:, 0].contiguous(), x[:, 1].contiguous(), x[:, 2].contiguous()
        ssim = M.SSIM()
        grid = (torch.rand(N, H, W, 2, device=dev) * 2 - 1)
        go = torch.rand(N, C, H, W, device=dev)
        g1 = torch.rand(N, 1, H, W, device=dev)
        cases = []

        def bwd_case(make_out, leaves, gout):
            def run():
                for t in leaves:
                    t.grad = None
                make_out().backward(gout)
            return run

        ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        cases.append(("ssim fwd", lambda: ssim(a, b), 3 * px))
        cases.append(("ssim fwd+bwd", bwd_case(lambda: ssim(ar, br), [ar, br], go), 3 * px + 5 * px))
        cr = c2.clone().requires_grad_(True)
        cases.append(("prediction_loss fwd (S=2)", lambda: M.prediction_loss(ssim, [a, c2], b), 3 * px + 2 * p1))
        cases.append(("prediction_loss fwd+bwd (S=2)", bwd_case(lambda: M.prediction_loss(ssim, [ar, cr], br), [ar, cr, br], g1), 2 * (3 * px + 2 * p1) + 3 * px))
        cases.append(("automasking_loss fwd (frame views)", lambda: M.automasking_loss(ssim, x, x[:, 1], (0, 2)), 3 * px + p1))
        gr = grid.clone().requires_grad_(True)
        cases.append(("grid_sample(border) fwd", lambda: M.grid_sample(a, grid, padding_mode="border"), 2 * px + 2 * p1))
        cases.append(("grid_sample(border) fwd+bwd", bwd_case(lambda: M.grid_sample(ar, gr, padding_mode="border"), [ar, gr], go), 2 * px + 2 * p1 + 3 * px + 4 * p1))
        d_full = disps[-1]
        dr = d_full.clone().requires_grad_(True)
        cases.append(("smooth_loss fwd+bwd", bwd_case(lambda: M.smooth_loss(dr[:, 0], b), [dr], torch.ones((), device=dev)), 2 * (p1 + px) + p1))
        dd = [d.clone().requires_grad_(True) for d in disps]
        rr = [r.clone().requires_grad_(True) for r in rv]
        tt = [t.clone().requires_grad_(True) for t in tv]
        ab_f, _ = SY.algorithmic_bytes(W, H, N, C, 2, 4, g=0)
        ab_fb, _ = SY.algorithmic_bytes(W, H, N, C, 2, 4, g=1)
        unit = W * H * N * 4
        fwd_bytes = 4 * (1 + C + 2 * C) * unit
        with torch.no_grad():
            cases.append(("fused view_synthesis_loss fwd (4 scales)", lambda: M.view_synthesis_loss(x, disps, rv, tv, K, invK), fwd_bytes))
        xr = x.clone().requires_grad_(True)

        def fused_fb():
            for t in dd + rr + tt + [xr]:
                t.grad = None
            M.view_synthesis_loss(xr, dd, rr, tt, K, invK).backward()
        cases.append(("fused view_synthesis_loss fwd+bwd (4 scales, g=1)", fused_fb, ab_fb))
        for name, fn, nbytes in cases:
            if name.startswith("fused") and "fwd (" in name:
                with torch.no_grad():
                    warm, cold = timeit(fn)
            else:
                warm, cold = timeit(fn)
            print(json.dumps({"op": name, "W": W, "H": H, "N": N, "C": C, "warm_ms": round(warm, 4), "cold_ms": round(cold, 4), "MB": round(nbytes / 1e6, 2),
                              "cold_GBps": round(nbytes / (cold * 1e-3) / 1e9, 1), "cold_frac_of_hbm_peak": round(nbytes / (cold * 1e-3) / 1e9 / peak, 4),
                              "note": "through the Python autograd mirror (includes its allocations / torch glue)"}), flush=True)


if __name__ == "__main__":
    main()
