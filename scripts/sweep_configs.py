#!/usr/bin/env python
"""Warp+SSIM loss micro-benchmark over BASELINE.json's other configurations (configs[2..4]): the fused
forward+backward call (md2_view_synthesis_loss_fwdbwd, device buffers) at 416x128 .. 1024x320, several batch
sizes, C in {1,3}, automask on/off.  Prints one JSON line per case: step time (CUDA events, inputs rotated
through a ring larger than the L2), algorithmic GB/s (BASELINE.md section 3 work model) and the fraction of the
measured HBM peak.  Usage: python scripts/sweep_configs.py [--quick]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import monodepth2_jl_b200 as M  # noqa: E402
from monodepth2_jl_b200 import _lib as L  # noqa: E402
from monodepth2_jl_b200 import synthetic as SY  # noqa: E402

SCALES = (0.125, 0.25, 0.5, 1.0)


def run_case(W, H, N, Cc, automask, steps=200):
    dev = torch.device("cuda", 0)
    ctx = M.Context.get(dev)
    nb = min(N, 2)
    x, disps, rv, tv = SY.synthetic_batch(nb, Cc, H, W, seed=7)
    rep = (N + nb - 1) // nb
    tile = lambda t: t.repeat(rep, *([1] * (t.dim() - 1)))[:N].contiguous()
    x, disps, rv, tv = tile(x), [tile(d) for d in disps], [tile(r) for r in rv], [tile(t) for t in tv]
    K, invK = SY.make_K(W, H)
    K_cm, invK_cm = K.t().contiguous().to(dev), invK.t().contiguous().to(dev)
    abytes, per_unit = SY.algorithmic_bytes(W, H, N, Cc, 2, len(SCALES), m=1 if automask else 0, g=1)
    set_bytes = 4 * (2 * x.numel() + 2 * sum(d.numel() for d in disps))
    n_sets = max(2, min(24, int(2.5 * 126e6 / set_bytes) + 1))
    sets = []
    for i in range(n_sets):
        xs = (x + 0.01 * i).clamp(0, 1).to(dev)
        auto = M.automasking_loss(M.SSIM(), xs, xs[:, 1], (0, 2)).contiguous() if automask else None
        st = dict(x=xs, disps=[d.to(dev) for d in disps], rv=[r.to(dev) for r in rv], tv=[t.to(dev) for t in tv],
                  loss=torch.zeros((), device=dev), gd=[torch.empty_like(d, device=dev) for d in disps],
                  gr=[torch.empty(N, 3, device=dev) for _ in range(2)], gt=[torch.empty(N, 3, device=dev) for _ in range(2)],
                  gx=torch.zeros(x.shape, device=dev), auto=auto)
        st["desc"] = L.make_vsl_desc(
            target=xs[:, 1], target_stride=xs.stride(0), sources=[xs[:, 0], xs[:, 2]], source_strides=[xs.stride(0)] * 2,
            disparities=st["disps"], K_cm=K_cm, invK_cm=invK_cm, rot=st["rv"], trans=st["tv"], pose_mode=1, invert=[1, 0],
            automask=auto, smooth_weight=[1e-3 * s for s in SCALES], loss_scale=0.25, normalize_disparity=True, loss=st["loss"],
            grad_disparity=st["gd"], grad_rot=st["gr"], grad_trans=st["gt"], grad_source=[st["gx"][:, 0], st["gx"][:, 2]],
            zero_grad_source=True, shape=(N, Cc, H, W))
        sets.append(st)
    stream = torch.cuda.current_stream(dev)
    f = ctx.lib.md2_view_synthesis_loss_fwdbwd

    def step(i):
        if f(ctx.handle, C.byref(sets[i % n_sets]["desc"]), 1.0, stream.cuda_stream):
            raise RuntimeError(ctx.lib.md2_last_error().decode())
    for i in range(max(3, n_sets)):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        step(i)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ctx.profile(True)
    for i in range(steps):
        step(i)
    kms, kn = ctx.profile_read()
    ctx.profile(False)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    k_ms = kms / max(kn, 1)
    return {"W": W, "H": H, "N": N, "C": Cc, "automask": bool(automask), "ms_per_step": round(ms, 4), "frames_per_s": round(N / ms * 1e3, 1),
            "march_kernel_ms": round(k_ms, 4), "bytes_per_unit": per_unit, "algorithmic_MB": round(abytes / 1e6, 1),
            "step_GBps": round(abytes / ms / 1e6, 1), "kernel_GBps": round(abytes / k_ms / 1e6, 1), "kernel_frac_of_hbm_peak": round(abytes / k_ms / 1e6 / peak, 4),
            "l2": f"ring of {n_sets} sets ({n_sets * set_bytes / 1e6:.0f} MB)"}


def main():
    quick = "--quick" in sys.argv
    cases = [(416, 128, 1, 1, 0), (416, 128, 8, 1, 0), (416, 128, 8, 3, 0), (416, 128, 8, 3, 1), (416, 128, 32, 1, 0), (416, 128, 64, 3, 1),
             (640, 192, 12, 3, 1), (640, 192, 12, 1, 0), (1024, 320, 4, 3, 0), (1024, 320, 4, 3, 1), (1024, 320, 16, 3, 1)]
    if quick:
        cases = cases[:3]
    if "--c3" in sys.argv:
        cases = [(416, 128, 8, 3, 0), (640, 192, 12, 3, 1), (1024, 320, 4, 3, 1)]
    for c in cases:
        print(json.dumps(run_case(*c)), flush=True)


if __name__ == "__main__":
    main()
