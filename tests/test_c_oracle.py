"""The second, independent oracle (oracle/c_oracle.c: plain-C float64 loops with a hand-derived reverse pass, SURVEY.md
section 7 step 1b) against (a) the known-answer vectors of the reference's test/runtests.jl, (b) the first oracle
(oracle/torch_oracle.py: torch primitives + autograd) on seeded inputs, value and every gradient element, (c) central
finite differences of its own value.  Two restatements that share nothing but the reference's formulas and agree to
float64 rounding pin each other; the -m gpu half (tests/test_gpu_c_oracle.py) closes the triangle with the CUDA path."""
import math

import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import torch_oracle as O
from util import oracle_vsl, rel_max

F64 = torch.float64


def _rodrigues(v):
    from scipy.spatial.transform import Rotation
    return torch.tensor(Rotation.from_rotvec(np.asarray(v)).as_matrix(), dtype=F64)


# ------------------------------- (a) test/runtests.jl known answers -------------------------------

def test_rotations():  # test/runtests.jl:14-29
    v = torch.rand(3, 3, dtype=F64)
    R = CO.so3_exp_map(v)
    for i in range(3):
        assert torch.allclose(R[i], _rodrigues(v[i].numpy()), atol=1e-5)


def test_transformation():  # test/runtests.jl:31-50
    rvec, tvec, p = torch.rand(1, 3, dtype=F64), torch.rand(1, 3, dtype=F64), torch.rand(3, dtype=F64)
    R, t = CO.composeT(rvec, tvec, False)
    np_ = R[0] @ p + t[0]
    assert torch.allclose(np_, _rodrigues(rvec[0].numpy()) @ p + tvec[0], atol=1e-6)
    Ri, ti = CO.composeT(rvec, tvec, True)
    assert torch.allclose(Ri[0] @ np_ + ti[0], p, atol=1e-6)


def test_ssim():  # test/runtests.jl:52-68
    one = torch.ones(1, 1, 2, 2, dtype=F64)
    assert torch.equal(CO.ssim(one, one), torch.zeros_like(one))
    assert torch.allclose(CO.ssim(one, torch.zeros_like(one)), torch.full_like(one, 0.49995000499950004), atol=1e-14)
    a, b = torch.rand(2, 1, 2, 2, dtype=F64), torch.rand(2, 1, 2, 2, dtype=F64)
    assert torch.allclose(CO.ssim(a, b), CO.ssim(b, a), atol=1e-15)


def test_smooth_loss():  # test/runtests.jl:70-83
    disp = torch.tensor([[0.0, 0.2], [0.1, 0.3]], dtype=F64).reshape(1, 2, 2)
    assert abs(CO.smooth_loss(disp, torch.ones(1, 1, 2, 2, dtype=F64)) - 0.3) < 1e-12
    image = torch.tensor([[0.1, 0.3], [0.2, 0.4]], dtype=F64).reshape(1, 1, 2, 2)
    sl = CO.smooth_loss(disp, image)
    assert abs(sl - 0.2542) < 1e-4 and abs(sl - 0.25422989241919236) < 1e-12
    assert abs(sl - (0.2 * math.exp(-0.2) + 0.1 * math.exp(-0.1))) < 1e-12


def test_disparity_to_depth():  # test/runtests.jl:85-92
    depth = CO.disparity_to_depth(torch.rand(2, 32, 32, dtype=F64), 0.1, 100.0)
    assert depth.min() >= 0.1 and depth.max() <= 100.0


def test_identity_warp():  # test/runtests.jl:94-122, through the full loss: identical frames + identity pose
    res, N = 16, 2
    torch.manual_seed(1)
    img = torch.rand(N, 1, 1, res, res, dtype=F64).expand(N, 3, 1, res, res).contiguous()
    K = torch.tensor([[910.0, 0, res / 2], [0, 910.0, res / 2], [0, 0, 1]], dtype=F64)
    disp = torch.rand(N, 1, res, res, dtype=F64)
    z = [torch.zeros(N, 3, dtype=F64)] * 2
    out = CO.view_synthesis_loss(img, [disp], z, z, K, torch.linalg.inv(K), scales=(1.0,), grad=False, viz=True)
    for w in out["viz_warped"]:
        assert torch.allclose(w, img[:, 1], atol=1e-3)
    assert out["viz_loss"].abs().max() < 1e-3


def test_pose_derivative():  # test/runtests.jl:124-142 (values from SURVEY.md section 4)
    x, target = torch.tensor([3.0, 2.0, 1.0], dtype=F64), torch.tensor([1.0, 2.0, 3.0], dtype=F64)
    r = torch.tensor([[1.0, 0.0, 0.0]], dtype=F64)
    R = CO.so3_exp_map(r)
    y = R[0] @ x - target
    l = y.norm()
    assert abs(l.item() - 2.775608012559207) < 1e-12
    yb = y / l                                       # d l / d (R x + t)
    gr = CO.so3_exp_map_bwd(r, torch.outer(yb, x).reshape(1, 3, 3))
    assert torch.allclose(gr[0], torch.tensor([1.3435210063, 1.1003665905, -2.8688708364], dtype=F64), atol=1e-9)
    assert torch.allclose(yb, torch.tensor([0.7205628428, -0.6344074398, -0.2798506565], dtype=F64), atol=1e-9)


def test_so3_below_the_clamp():  # max.(theta, 1e-4) (src/utils.jl:110): derivative of the clamp is 0 below it
    r = torch.tensor([[3e-5, -2e-5, 1e-5]], dtype=F64, requires_grad=True)
    d = torch.rand(1, 3, 3, dtype=F64)
    (g,) = torch.autograd.grad((O.so3_exp_map(r) * d).sum(), r)
    assert torch.allclose(CO.so3_exp_map_bwd(r, d), g, rtol=1e-10, atol=1e-14)


# ------------------------------- (b) the two oracles against each other -------------------------------

def test_primitives_agree():
    torch.manual_seed(3)
    x, y = torch.rand(2, 3, 7, 9, dtype=F64), torch.rand(2, 3, 7, 9, dtype=F64)
    assert torch.allclose(CO.ssim(x, y), O.SSIM()(x, y), atol=1e-14)
    d = torch.rand(2, 7, 9, dtype=F64, requires_grad=True)
    v, g = CO.smooth_loss(d, x, grad=True)
    ref = O.smooth_loss(d, x)
    (gr,) = torch.autograd.grad(ref, d)
    assert abs(v - ref.item()) < 1e-14 and torch.allclose(g, gr, atol=1e-15)
    lo = torch.rand(2, 1, 3, 4, dtype=F64)
    assert torch.allclose(CO.upsample_bilinear(lo, (9, 7)), O.upsample_bilinear(lo, (9, 7)), atol=1e-14)
    grid = torch.rand(2, 7, 9, 2, dtype=F64) * 2.6 - 1.3             # a third of the points outside the image
    assert torch.allclose(CO.grid_sample_border(x, grid), O.grid_sample(x, grid, padding_mode="border"), atol=1e-14)
    rv, tv = torch.randn(4, 3, dtype=F64), torch.randn(4, 3, dtype=F64)
    for inv in (False, True):
        (Ra, ta), (Rb, tb) = CO.composeT(rv, tv, inv), O.composeT(rv, tv, inv)
        assert torch.allclose(Ra, Rb, atol=1e-14) and torch.allclose(ta, tb, atol=1e-14)


def _compare(out, ref, tol=1e-9, ids=(0, 2)):
    assert abs(out["loss"] - ref["loss"]) <= 1e-12 * abs(ref["loss"]), (out["loss"], ref["loss"])
    for k in ("gdisp", "grvec", "gtvec"):
        for i, (a, b) in enumerate(zip(out[k], ref[k])):
            assert rel_max(a, b) <= tol, (k, i, rel_max(a, b))
    a, b = out["gx"][:, list(ids)], ref["gx"][:, list(ids)]
    assert rel_max(a, b) <= tol, ("gx", rel_max(a, b))


@pytest.mark.parametrize("N,C,H,W,automask,seed", [(2, 3, 24, 40, True, 5), (1, 1, 17, 23, False, 6), (2, 1, 32, 48, True, 7),
                                                   (1, 3, 2, 2, False, 8)])
def test_full_loss_and_every_gradient_element(N, C, H, W, automask, seed):
    """float64 against float64: the same piece of the piecewise-smooth loss (decisions agree unless a margin is within
    float64 rounding), so every gradient element has to agree to ~1e-9 relative."""
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=seed, dtype=F64, pose_sigma=0.02)
    K, invK = O.make_K(W, H, dtype=F64)
    ref = oracle_vsl(x, disps, rv, tv, K, invK, automask=automask)
    out = CO.view_synthesis_loss(x, disps, rv, tv, K, invK, auto_loss=ref["auto"])
    _compare(out, ref)


def test_other_frames_scales_and_sources():
    """one source that is not frame 0, target frame 0 (no inverted pose), two decoder scales, other depth range / weights"""
    x, disps, rv, tv = O.synthetic_batch(2, 3, 20, 28, seed=9, dtype=F64, pose_sigma=0.03)
    K, invK = O.make_K(28, 20, dtype=F64)
    kw = dict(target_id=0, source_ids=(2,), scales=(0.5, 1.0), min_depth=0.5, max_depth=50.0, disparity_smoothness=3e-2)
    ref = oracle_vsl(x, disps[2:], rv[:1], tv[:1], K, invK, **kw)
    out = CO.view_synthesis_loss(x, disps[2:], rv[:1], tv[:1], K, invK, **kw)
    _compare(out, ref, ids=(2,))


def test_points_leaving_the_image_and_behind_the_camera():
    """large poses: border clipping (gradient masks) and c3 <= 0"""
    x, disps, rv, tv = O.synthetic_batch(1, 1, 16, 24, seed=10, dtype=F64, pose_sigma=0.3)
    tv[0][0, 2] = 150.0                                            # beyond max_depth: every point behind the camera
    K, invK = O.make_K(24, 16, dtype=F64)
    ref = oracle_vsl(x, disps, rv, tv, K, invK)
    out = CO.view_synthesis_loss(x, disps, rv, tv, K, invK)
    _compare(out, ref)


def test_visualisation_outputs():  # src/training.jl:34-37, 71-74
    x, disps, rv, tv = O.synthetic_batch(2, 3, 12, 20, seed=11, dtype=F64)
    K, invK = O.make_K(20, 12, dtype=F64)
    loss, (warped, wl) = O.view_synthesis_loss(x, disps, rv, tv, K, invK, return_viz=True)
    out = CO.view_synthesis_loss(x, disps, rv, tv, K, invK, grad=False, viz=True)
    assert abs(out["loss"] - loss.item()) < 1e-13
    assert torch.allclose(out["viz_loss"], wl, atol=1e-13)
    for a, b in zip(out["viz_warped"], warped):
        assert torch.allclose(a, b, atol=1e-13)


def test_simple_depth_objective():  # src/simple_depth.jl:25-41 (config 1)
    x, _, rv, tv = O.synthetic_batch(1, 3, 16, 24, seed=12, dtype=F64)
    K, invK = O.make_K(24, 16, dtype=F64)
    disp = (torch.rand(1, 1, 16, 24, dtype=F64) * 0.5 + 0.25).requires_grad_(True)
    rq = [r.clone().requires_grad_(True) for r in rv]
    tq = [t.clone().requires_grad_(True) for t in tv]
    loss = O.simple_depth_loss(x, disp, rq, tq, K, invK)
    g = torch.autograd.grad(loss, [disp] + rq + tq)
    out = CO.simple_depth_loss(x, disp, rv, tv, K, invK)
    assert abs(out["loss"] - loss.item()) < 1e-13
    assert rel_max(out["gdisp"][0], g[0]) < 1e-9
    for s in range(2):
        assert rel_max(out["grvec"][s], g[1 + s]) < 1e-9 and rel_max(out["gtvec"][s], g[3 + s]) < 1e-9


def test_min_tie_routing():  # the mask wins a tie with the warp loss (it is first in the cat)
    x = torch.rand(1, 1, 1, 6, 6, dtype=F64).expand(1, 3, 1, 6, 6).contiguous()     # identical frames: pe = 0 for both sources
    K, invK = O.make_K(6, 6, dtype=F64)
    tiny = [torch.full((1, 3), 1e-300, dtype=F64)] * 2                               # (theta = 0 itself gives the reference's NaN)
    zero = [torch.zeros(1, 3, dtype=F64)] * 2
    disp = torch.full((1, 1, 6, 6), 0.5, dtype=F64)
    out = CO.view_synthesis_loss(x, [disp], tiny, zero, K, invK, scales=(1.0,), auto_loss=torch.zeros(1, 1, 6, 6, dtype=F64))
    assert out["loss"] == 0.0 and out["gx"].abs().max() == 0                         # everything routed to the (constant) mask
    # ... and source 0 wins a tie with source 1: both sources are the same frame seen through the same pose
    x2, disps, rv, tv = O.synthetic_batch(1, 1, 10, 12, seed=14, dtype=F64)
    x2[:, 2] = x2[:, 0]
    K, invK = O.make_K(12, 10, dtype=F64)
    x4 = torch.cat([x2, x2[:, :1]], 1)                                               # frames 2 and 3 are the same image
    out = CO.view_synthesis_loss(x4, disps, [rv[0]] * 2, [tv[0]] * 2, K, invK, target_id=1, source_ids=(2, 3))
    assert out["gx"][:, 2].abs().max() > 0 and out["gx"][:, 3].abs().max() == 0
    assert out["grvec"][1].abs().max() == 0 and out["gtvec"][1].abs().max() == 0


# ------------------------------- (c) finite differences of its own value -------------------------------

def test_gradients_match_central_differences():
    x, disps, rv, tv = O.synthetic_batch(1, 1, 12, 16, seed=13, dtype=F64, pose_sigma=0.02)
    K, invK = O.make_K(16, 12, dtype=F64)
    out = CO.view_synthesis_loss(x, disps, rv, tv, K, invK)
    g = torch.Generator().manual_seed(0)
    val = lambda dd, rr, tt, xx: CO.view_synthesis_loss(xx, dd, rr, tt, K, invK, grad=False)["loss"]
    eps = 1e-6
    for trial in range(3):
        vd = [torch.randn(d.shape, generator=g, dtype=F64) for d in disps]
        vr = [torch.randn(r.shape, generator=g, dtype=F64) for r in rv]
        vt = [torch.randn(t.shape, generator=g, dtype=F64) for t in tv]
        vx = torch.zeros_like(x)
        vx[:, [0, 2]] = torch.randn(x[:, [0, 2]].shape, generator=g, dtype=F64)
        step = lambda s: val([d + s * v for d, v in zip(disps, vd)], [r + s * v for r, v in zip(rv, vr)],
                             [t + s * v for t, v in zip(tv, vt)], x + s * vx)
        fd = (step(eps) - step(-eps)) / (2 * eps)
        an = sum((a * b).sum() for a, b in zip(out["gdisp"], vd)) + sum((a * b).sum() for a, b in zip(out["grvec"], vr)) + \
            sum((a * b).sum() for a, b in zip(out["gtvec"], vt)) + (out["gx"] * vx).sum()
        assert abs(fd - an.item()) <= 2e-5 * max(abs(fd), 1e-12) + 1e-9, (trial, fd, an.item())


# ------------------------------- forced decisions (the flip-controlled strict parity of tests/test_*_forced.py) -------------------------------

@pytest.mark.parametrize("N,C,H,W,am,seed", [(2, 1, 32, 64, False, 3), (1, 3, 30, 37, True, 6)])
def test_forced_decisions_three_way(N, C, H, W, am, seed):
    """The decisions of the float32 marching-kernel source (run by tests/emul) forced into BOTH float64 oracles: they agree
    with each other to rounding, and the float32 implementation with the C oracle within BASELINE.json's bars."""
    from emul_util import emul_vsl
    from util import check_vsl, oracle_vsl_forced
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=seed)
    K, invK = O.make_K(W, H)
    auto = O.automasking_loss(O.SSIM(), x.double(), x.double()[:, 1], (0, 2)) if am else None
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, automask=auto.float().contiguous() if am else None, debug_choices=True, R=12)
    out["loss"] = out["loss"].item()
    ref_t = oracle_vsl_forced(x, disps, rv, tv, K, invK, out["choices"], auto=auto)
    ref_c = CO.view_synthesis_loss(x, disps, rv, tv, K, invK, auto_loss=auto, choices=out["choices"])
    _compare(ref_c, ref_t)
    check_vsl(out, ref_c, tag=f"emul vs C oracle, forced {N},{C},{H},{W},{am}")


# ------------------------------- the committed golden fixtures (tests/golden/, made by the first oracle) -------------------------------

def test_golden_vsl_small():
    """the C oracle reproduces the float64 loss / gradients the torch oracle committed for the seeded 4-scale automask case"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vsl_small.npz"))
    x = torch.from_numpy(g["x"])
    disps = [torch.from_numpy(g[f"disp{i}"]) for i in range(4)]
    rv = [torch.from_numpy(g[f"rvec{s}"]) for s in range(2)]
    tv = [torch.from_numpy(g[f"tvec{s}"]) for s in range(2)]
    # (the fixture keeps the automask map in float32; the committed loss was formed with the float64 map)
    auto = O.automasking_loss(O.SSIM(), x.double(), x.double()[:, 1], (0, 2))
    assert torch.allclose(auto.float(), torch.from_numpy(g["auto"]), atol=1e-7)
    out = CO.view_synthesis_loss(x, disps, rv, tv, torch.from_numpy(g["K"]), torch.from_numpy(g["invK"]), auto_loss=auto)
    assert abs(out["loss"] - float(g["loss"])) < 1e-12
    for i in range(4):
        assert rel_max(out["gdisp"][i], torch.from_numpy(g[f"gdisp{i}"])) < 1e-9
    for s in range(2):
        assert rel_max(out["grvec"][s], torch.from_numpy(g[f"grvec{s}"])) < 1e-9
        assert rel_max(out["gtvec"][s], torch.from_numpy(g[f"gtvec{s}"])) < 1e-9


def test_golden_simple_depth_c1():
    """config 1 on the reference's own res/image.png triplet (src/dtk.jl:16-47): the slow_depth start point and a second point"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simple_depth_c1.npz"))
    x = torch.from_numpy(g["frames"]).permute(0, 3, 1, 2).to(F64).div(255.0).unsqueeze(0).contiguous()
    W, H = 416, 128
    K, invK = O.make_K(W, H, f=float(g["focal"]), dtype=F64)
    disp = torch.full((1, 1, H, W), 0.5, dtype=F64)
    rv = [torch.tensor([[0.0, 0.0, 0.01]], dtype=F64) for _ in range(2)]
    tv = [torch.zeros(1, 3, dtype=F64) for _ in range(2)]
    out = CO.simple_depth_loss(x, disp, rv, tv, K, invK)
    assert abs(out["loss"] - float(g["loss"])) < 1e-12
    for s in range(2):
        assert rel_max(out["grvec"][s], torch.from_numpy(g["grvec"][s])) < 1e-9
        assert rel_max(out["gtvec"][s], torch.from_numpy(g["gtvec"][s])) < 1e-9
    yy, xx = torch.meshgrid(torch.arange(H, dtype=F64), torch.arange(W, dtype=F64), indexing="ij")
    disp2 = (0.5 + 0.2 * torch.sin(xx / 37.0) * torch.cos(yy / 23.0)).reshape(1, 1, H, W)
    rv2 = [torch.from_numpy(g["rvec2"][s]) for s in range(2)]
    tv2 = [torch.from_numpy(g["tvec2"][s]) for s in range(2)]
    out = CO.simple_depth_loss(x, disp2, rv2, tv2, K, invK)
    assert abs(out["loss"] - float(g["loss2"])) < 1e-12
    assert rel_max(out["gdisp"][0], torch.from_numpy(g["gdisp2"]).double()) < 1e-6      # (committed as float32)
    for s in range(2):
        assert rel_max(out["grvec"][s], torch.from_numpy(g["grvec2"][s])) < 1e-9
        assert rel_max(out["gtvec"][s], torch.from_numpy(g["gtvec2"][s])) < 1e-9


@pytest.mark.parametrize("trial", range(8))
def test_fuzz_marching_source_against_the_c_oracle(trial):
    """random odd shapes / scale counts / chunk heights: the float32 marching-kernel source (tests/emul) against the C oracle with
    its decisions forced -- EVERY gradient element within BASELINE.json's bars.  (The two share nothing: a bug common to the
    kernel's formulation and the torch oracle's primitives would show here.)"""
    import random
    from emul_util import emul_vsl
    from util import check_vsl
    rnd = random.Random(1000 + trial)
    N, C = rnd.choice([1, 2, 3]), rnd.choice([1, 3])
    H, W = rnd.randint(2, 70), rnd.randint(2, 120)
    L, am, R = rnd.choice([1, 2, 3, 4]), rnd.random() < 0.5, rnd.choice([8, 12, 16, 33])
    scales = (0.125, 0.25, 0.5, 1.0)[4 - L:]
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, scales=scales, seed=200 + trial, pose_sigma=rnd.choice([0.01, 0.03]))
    K, invK = O.make_K(W, H)
    auto = O.automasking_loss(O.SSIM(), x.double(), x.double()[:, 1], (0, 2)) if am else None
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, scales=scales, automask=auto.float().contiguous() if am else None,
                   debug_choices=True, R=R)
    out["loss"] = out["loss"].item()
    ref = CO.view_synthesis_loss(x, disps, rv, tv, K, invK, scales=scales, auto_loss=auto, choices=out["choices"])
    check_vsl(out, ref, tag=f"fuzz {trial}: N={N} C={C} {W}x{H} L={L} automask={am} R={R}")


def test_config2_full_size_marching_source_against_the_c_oracle():
    """BASELINE.json configs[1] at full size (416x128, batch 8, C = 1, 4 native-size scales; the very batch bench.py times, chunk
    height 33 as chosen on the B200): the marching-kernel source under the fibre emulator against the C oracle with its
    decisions forced -- loss 1e-5, EVERY gradient element 1e-4."""
    from emul_util import emul_vsl
    from util import check_vsl
    x, disps, rv, tv = O.synthetic_batch(8, 1, 128, 416, seed=42)
    K, invK = O.make_K(416, 128)
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, debug_choices=True, R=33)
    out["loss"] = out["loss"].item()
    ref = CO.view_synthesis_loss(x, disps, rv, tv, K, invK, choices=out["choices"])
    check_vsl(out, ref, tag="config 2, full size")
    # ... and the decisions themselves: the float64 oracle, left to decide on its own, takes the same branch at all but a
    # handful of the 1.7 M pixels (forcing cannot hide a wrong cell / arg-min / mask: it would show here)
    own = CO.view_synthesis_loss(x, disps, rv, tv, K, invK, grad=False, export_choices=True)
    assert abs(own["loss"] - ref["loss"]) <= 1e-6 * abs(ref["loss"])
    from util import DECISION_FLIP_MAX, decision_mismatch
    for kind, f in decision_mismatch(out["choices"], own["choices"], 1).items():
        assert f <= DECISION_FLIP_MAX, (kind, f)


def test_config3_geometry_marching_source_against_the_c_oracle():
    """BASELINE.json configs[2] geometry (640x192, C = 3, automask + min-reprojection), 2 of its 12 images (the emulator runs
    every warp as 32 fibres: the whole batch would take minutes; all 12 run on the GPU in tests/test_gpu_forced.py)"""
    from emul_util import emul_vsl
    from util import check_vsl
    x, disps, rv, tv = O.synthetic_batch(2, 3, 192, 640, seed=42)
    K, invK = O.make_K(640, 192)
    auto = O.automasking_loss(O.SSIM(), x.double(), x.double()[:, 1], (0, 2))
    out = emul_vsl(x, disps, rv, tv, K, invK, mode=2, automask=auto.float().contiguous(), debug_choices=True, R=32)
    out["loss"] = out["loss"].item()
    ref = CO.view_synthesis_loss(x, disps, rv, tv, K, invK, auto_loss=auto, choices=out["choices"])
    check_vsl(out, ref, tag="config 3 geometry")
    own = CO.view_synthesis_loss(x, disps, rv, tv, K, invK, auto_loss=auto, grad=False, export_choices=True)
    from util import DECISION_FLIP_MAX, decision_mismatch
    for kind, f in decision_mismatch(out["choices"], own["choices"], 3).items():
        assert f <= DECISION_FLIP_MAX, (kind, f)
