"""CPU check of the fused CUDA kernel's tile logic: the __host__ __device__ phases of
csrc/md2_fused.cuh are run sequentially by tests/emul and compared with the float64 oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from emul_util import emul_vsl
from util import check_vsl, check_vsl_statistical, oracle_vsl, rel_l2, rel_max, well_conditioned_batch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("N,C,H,W,am", [(1, 1, 24, 40, False), (1, 3, 24, 40, True), (1, 3, 20, 37, True),
                                        (1, 1, 17, 33, False)])
def test_fused_strict_on_well_conditioned_inputs(N, C, H, W, am):
    (x, disps, rv, tv, K, invK), seed = well_conditioned_batch(N, C, H, W, am)
    ref = oracle_vsl(x, disps, rv, tv, K, invK, automask=am)
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, automask=ref["auto"].float().contiguous() if am else None)
    out["loss"] = out["loss"].item()
    check_vsl(out, ref, tag=f"{N},{C},{H},{W},{am},seed={seed}")


@pytest.mark.parametrize("N,C,H,W,am", [(2, 1, 32, 64, False), (1, 3, 48, 80, True), (2, 3, 40, 100, True)])
def test_fused_statistical_on_arbitrary_inputs(N, C, H, W, am):
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=3)
    K, invK = O.make_K(W, H)
    ref = oracle_vsl(x, disps, rv, tv, K, invK, automask=am)
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, automask=ref["auto"].float().contiguous() if am else None)
    out["loss"] = out["loss"].item()
    check_vsl_statistical(out, ref, tag=f"{N},{C},{H},{W},{am}")


@pytest.mark.parametrize("R", [8, 32])
def test_stress_poses_strict(R):
    (x, disps, rv, tv, K, invK), seed = well_conditioned_batch(1, 3, 24, 40, True, pose_sigma=0.1)
    ref = oracle_vsl(x, disps, rv, tv, K, invK, automask=True)
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, automask=ref["auto"].float().contiguous(), R=R)
    out["loss"] = out["loss"].item()
    check_vsl(out, ref, tag=f"stress strict seed={seed}")


@pytest.mark.parametrize("R", [4, 12, 64])
def test_chunk_height_does_not_change_results(R):
    """rows per chunk only re-partitions the work: same loss / gradients for any R"""
    x, disps, rv, tv = O.synthetic_batch(1, 1, 40, 70, seed=9)
    K, invK = O.make_K(70, 40)
    a = emul_vsl(x, disps, rv, tv, K, invK, mode=2, R=R)
    b = emul_vsl(x, disps, rv, tv, K, invK, mode=2, R=32)
    assert abs(a["loss"].item() - b["loss"].item()) <= 2e-6 * abs(b["loss"].item())
    for u, v in zip(a["gdisp"] + a["grvec"] + a["gtvec"] + [a["gx"]], b["gdisp"] + b["grvec"] + b["gtvec"] + [b["gx"]]):
        assert rel_l2(u, v) < 2e-2     # (a tie flip would show up as ~1e-2 on a 70x40 image; rounding alone is ~1e-6)


def test_fwd_then_bwd_equals_fused():
    x, disps, rv, tv = O.synthetic_batch(1, 3, 32, 64, seed=5)
    K, invK = O.make_K(64, 32)
    fused = emul_vsl(x, disps, rv, tv, K, invK, mode=2)
    fwd = emul_vsl(x, disps, rv, tv, K, invK, mode=0)
    bwd = emul_vsl(x, disps, rv, tv, K, invK, mode=1, saved=fwd["saved"])
    assert abs(fwd["loss"].item() - fused["loss"].item()) <= 1e-6 * abs(fused["loss"].item())
    for a, b in zip(bwd["gdisp"], fused["gdisp"]):
        assert rel_max(a, b) < 1e-5
    for a, b in zip(bwd["grvec"] + bwd["gtvec"], fused["grvec"] + fused["gtvec"]):
        assert rel_max(a, b) < 1e-5
    # linearity in the upstream cotangent
    half = emul_vsl(x, disps, rv, tv, K, invK, mode=1, saved=fwd["saved"], gloss=0.5)
    assert rel_max(half["gdisp"][3] * 2, bwd["gdisp"][3]) < 1e-6


def test_tiny_2x2_and_viz():
    # reflect padding needs W,H >= 2; the reference's own tests use 2x2 inputs
    torch.manual_seed(0)
    x = torch.rand(1, 3, 1, 2, 2)
    disps = [torch.rand(1, 1, 2, 2) * 0.5 + 0.2]
    rv = [torch.tensor([[0.0, 0.0, 0.01]]), torch.tensor([[0.01, 0.0, 0.0]])]
    tv = [torch.zeros(1, 3), torch.zeros(1, 3)]
    K, invK = O.make_K(2, 2, f=5.0)
    ref = oracle_vsl(x, disps, rv, tv, K, invK, scales=(1.0,))
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, scales=(1.0,), viz=True)
    assert abs(out["loss"].item() - ref["loss"]) < 1e-5 * abs(ref["loss"])
    xd = x.double()
    loss, (warped, wl) = O.view_synthesis_loss(xd, [d.double() for d in disps], [r.double() for r in rv],
                                               [t.double() for t in tv], K.double(), invK.double(), scales=(1.0,),
                                               return_viz=True)
    assert torch.allclose(out["viz_loss"].double(), wl, atol=1e-5)
    for a, b in zip(out["viz_warped"], warped):
        assert torch.allclose(a.double(), b, atol=1e-5)


def test_golden_simple_depth_c1():
    """config 1: the reference's res/image.png triplet at the slow_depth start point"""
    g = np.load(os.path.join(GOLD, "simple_depth_c1.npz"))
    x = torch.from_numpy(g["frames"]).permute(0, 3, 1, 2).float().div(255.0).unsqueeze(0).contiguous()
    W, H = 416, 128
    K, invK = O.make_K(W, H, f=float(g["focal"]))
    disp = torch.full((1, 1, H, W), 0.5)
    rv = [torch.tensor([[0.0, 0.0, 0.01]]) for _ in range(2)]
    tv = [torch.zeros(1, 3) for _ in range(2)]
    out = emul_vsl(x, [disp], rv, tv, K, invK, mode=2, normalize=False, smooth_weight=[1.0], loss_scale=1.0)
    assert abs(out["loss"].item() - float(g["loss"])) <= 1e-5 * float(g["loss"])
    # t = 0: the projection does not depend on depth, so d loss / d disp is exactly 0 (noise only)
    assert out["gdisp"][0].abs().max() < 1e-7 and np.abs(g["gdisp"]).max() < 1e-7
    for s in range(2):
        assert rel_max(out["grvec"][s], torch.from_numpy(g["grvec"][s])) < 2e-3
        assert rel_max(out["gtvec"][s], torch.from_numpy(g["gtvec"][s])) < 2e-3
    # second point: smooth disparity bump + small translations
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    disp2 = (0.5 + 0.2 * torch.sin(xx / 37.0) * torch.cos(yy / 23.0)).reshape(1, 1, H, W).float()
    rv2 = [torch.from_numpy(g["rvec2"][s]).float() for s in range(2)]
    tv2 = [torch.from_numpy(g["tvec2"][s]).float() for s in range(2)]
    out = emul_vsl(x, [disp2], rv2, tv2, K, invK, mode=2, normalize=False, smooth_weight=[1.0], loss_scale=1.0)
    ref = dict(loss=float(g["loss2"]), gdisp=[torch.from_numpy(g["gdisp2"])],
               grvec=[torch.from_numpy(g["grvec2"][s]) for s in range(2)],
               gtvec=[torch.from_numpy(g["gtvec2"][s]) for s in range(2)])
    out["loss"] = out["loss"].item()
    out["gx"] = None
    check_vsl_statistical(out, ref, tag="golden c1 point 2")


def test_golden_vsl_small_pins_oracle():
    """the oracle reproduces its committed float64 outputs (guards against oracle drift)"""
    g = np.load(os.path.join(GOLD, "vsl_small.npz"))
    x = torch.from_numpy(g["x"])
    disps = [torch.from_numpy(g[f"disp{i}"]) for i in range(4)]
    rv = [torch.from_numpy(g[f"rvec{s}"]) for s in range(2)]
    tv = [torch.from_numpy(g[f"tvec{s}"]) for s in range(2)]
    K, invK = torch.from_numpy(g["K"]), torch.from_numpy(g["invK"])
    ref = oracle_vsl(x, disps, rv, tv, K, invK, automask=True)
    assert abs(ref["loss"] - float(g["loss"])) < 1e-12
    for i in range(4):
        assert torch.allclose(ref["gdisp"][i], torch.from_numpy(g[f"gdisp{i}"]), atol=1e-14)
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, automask=torch.from_numpy(g["auto"]))
    out["loss"] = out["loss"].item()
    check_vsl_statistical(out, ref, tag="golden vsl_small")


def test_closer_to_float64_than_float32_reference():
    """the fused kernel's float32 result is at least as close to the float64 truth as the
    reference's own op sequence evaluated in float32 (what its GPU path computes)"""
    N, C, H, W = 3, 1, 64, 200
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=3)
    K, invK = O.make_K(W, H)
    ref = oracle_vsl(x, disps, rv, tv, K, invK, automask=True)
    ref32 = oracle_vsl(x, disps, rv, tv, K, invK, automask=True, dtype=torch.float32)
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, automask=ref["auto"].float().contiguous())
    assert abs(out["loss"].item() - ref["loss"]) <= abs(ref32["loss"] - ref["loss"]) + 1e-9
    for l in range(4):
        assert rel_l2(out["gdisp"][l], ref["gdisp"][l]) <= rel_l2(ref32["gdisp"][l], ref["gdisp"][l]) + 1e-6
    for name in ("grvec", "gtvec"):
        for s in range(2):
            assert rel_max(out[name][s], ref[name][s]) <= rel_max(ref32[name][s], ref[name][s]) + 1e-6
