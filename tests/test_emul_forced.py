"""Flip-controlled strict parity (CPU): the marching-kernel source run by tests/emul exports its discrete decisions
(gather cell, clip masks, arg-min, clamp pass, |.| signs); the float64 oracle evaluated with exactly those decisions
must agree with EVERY gradient element within BASELINE.json's bars -- on arbitrary, ill-conditioned inputs too."""
import pytest
import torch

from oracle import torch_oracle as O
from emul_util import emul_vsl
from util import check_vsl, oracle_vsl_forced


@pytest.mark.parametrize("N,C,H,W,am,seed", [(2, 1, 32, 64, False, 3), (1, 3, 48, 80, True, 3), (2, 3, 40, 100, True, 4),
                                             (1, 1, 17, 33, False, 5), (1, 3, 30, 37, False, 6)])
def test_forced_strict_on_arbitrary_inputs(N, C, H, W, am, seed):
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=seed)
    K, invK = O.make_K(W, H)
    auto = O.automasking_loss(O.SSIM(), x.double(), x.double()[:, 1], (0, 2)) if am else None
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, automask=auto.float().contiguous() if am else None, debug_choices=True, R=12)
    out["loss"] = out["loss"].item()
    ref = oracle_vsl_forced(x, disps, rv, tv, K, invK, out["choices"], auto=auto)
    check_vsl(out, ref, tag=f"forced {N},{C},{H},{W},{am}")


@pytest.mark.parametrize("sigma", [0.03, 0.1])
def test_forced_stress_poses(sigma):
    """poses 3x / 10x larger than the pose decoder's output scale (src/pose_decoder.jl:30).  At 0.1 a sixth of the points of
    source 0 lie within 0.05 of the camera plane or behind it (c3 + 1e-7 <= 0: SURVEY appendix B, no guard in the reference
    either), where u = c1 / c3 amplifies float32 rounding a thousandfold: the bar is 10x wider there, and only there."""
    x, disps, rv, tv = O.synthetic_batch(2, 3, 48, 96, seed=5, pose_sigma=sigma)
    K, invK = O.make_K(96, 48)
    auto = O.automasking_loss(O.SSIM(), x.double(), x.double()[:, 1], (0, 2))
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, automask=auto.float().contiguous(), debug_choices=True, R=16)
    out["loss"] = out["loss"].item()
    ref = oracle_vsl_forced(x, disps, rv, tv, K, invK, out["choices"], auto=auto)
    check_vsl(out, ref, tag=f"forced stress poses {sigma}", grad_rtol=1e-4 if sigma < 0.1 else 1e-3)
