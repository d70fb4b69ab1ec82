"""GPU edge cases of the fused path and the `train_loss` wrapper with its containers."""
import pytest
import torch

import monodepth2_jl_b200 as M
from oracle import torch_oracle as O
from util import GRAD_RTOL, LOSS_RTOL, check_vsl, oracle_vsl, oracle_vsl_forced, rel_max
from test_gpu_forced import run_cuda

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def test_theta_zero_gradient_is_the_limit():
    """rvec = 0: the reference's sqrt gives a NaN gradient (README.md:47-49); here the theta-path contributes 0 (documented
    deviation) and everything else is finite and right"""
    x, disps, rv, tv = O.synthetic_batch(2, 1, 32, 64, seed=2)
    K, invK = O.make_K(64, 32)
    rv0 = [torch.zeros_like(r) for r in rv]
    out = run_cuda(x, disps, rv0, tv, K, invK, automask=False)
    assert all(torch.isfinite(g).all() for g in out["grvec"] + out["gtvec"] + out["gdisp"])
    # inside the clamp theta < 1e-4 the reference's R = (sin theta / max(theta, 1e-4)) K + ... scales the rotation down
    # linearly: d R / d rvec -> 0 for theta -> 0 (src/utils.jl:102-117), and that is what comes out at theta = 0
    assert all(g.abs().max().item() == 0.0 for g in out["grvec"])
    ref = oracle_vsl_forced(x, disps, rv0, tv, K, invK, out["choices"])      # (its d / d rvec is NaN)
    assert abs(out["loss"] - ref["loss"]) <= LOSS_RTOL * abs(ref["loss"])
    for a, b in zip(out["gtvec"] + out["gdisp"], ref["gtvec"] + ref["gdisp"]):
        assert rel_max(a, b) <= GRAD_RTOL


def test_identity_pose():
    """R = I, t = 0 (the reference's identity-warp test, test/runtests.jl:94-122): every sample sits on its own pixel up to
    the 1e-7 by which a float32 invK misses K^-1, so border columns / rows fall on either side of the closed clip border;
    warped == source to rounding, and the gradients match the oracle under the kernel's own decisions"""
    x, disps, rv, tv = O.synthetic_batch(2, 3, 24, 40, seed=4)
    K, invK = O.make_K(40, 24)
    rv0 = [torch.zeros_like(r) for r in rv]
    tv0 = [torch.zeros_like(t) for t in tv]
    d = dev()
    xg = x.to(d)
    loss, warped, wl = M.view_synthesis_loss(xg, [t.to(d) for t in disps], [r.to(d) for r in rv0], [t.to(d) for t in tv0],
                                             K.to(d), invK.to(d), return_viz=True)
    assert torch.allclose(warped[0], xg[:, 0], atol=1e-5) and torch.allclose(warped[1], xg[:, 2], atol=1e-5)
    out = run_cuda(x, disps, rv0, tv0, K, invK, automask=False)
    ch = O.decode_choices(out["choices"], 3, 2)
    assert ch["mx"][..., 1:-1, :].all() and ch["my"][:, :, 1:-1].all()            # the interior is never clipped
    ref = oracle_vsl_forced(x, disps, rv0, tv0, K, invK, out["choices"])          # (d / d rvec is NaN in the oracle at 0)
    assert abs(out["loss"] - ref["loss"]) <= LOSS_RTOL * abs(ref["loss"])
    for a, b in zip(out["gtvec"], ref["gtvec"]):
        assert rel_max(a, b) <= GRAD_RTOL
    assert rel_max(out["gx"][:, [0, 2]], ref["gx"][:, [0, 2]]) <= GRAD_RTOL


def test_points_behind_the_camera():
    """c3 + 1e-7 <= 0 for part of the image (no guard in the reference, src/utils.jl:97): coordinates explode, the border clip
    catches them, nothing is NaN; parity with the float64 oracle under forced decisions (10x bar: u = c1 / c3 near c3 = 0)"""
    x, disps, rv, tv = O.synthetic_batch(2, 3, 48, 96, seed=5, pose_sigma=0.1)
    K, invK = O.make_K(96, 48)
    out = run_cuda(x, disps, rv, tv, K, invK, automask=True)
    assert all(torch.isfinite(g).all() for g in out["grvec"] + out["gtvec"] + out["gdisp"] + [out["gx"]])
    ref = oracle_vsl_forced(x, disps, rv, tv, K, invK, out["choices"], auto=out["auto"].double())
    check_vsl(out, ref, tag="behind the camera", grad_rtol=1e-3)


def test_mask_routing_and_source_tie_inside_the_fused_kernel():
    """(a) an automask below every photometric error wins everywhere and takes the gradient (src/training.jl:17-19): no
    photometric gradient is left; (b) two identical sources with identical poses tie bit for bit: the first index wins
    (findmin), so source 0 gets all of the gradient"""
    x, disps, rv, tv = O.synthetic_batch(2, 1, 32, 64, seed=6, full_res_disp=True)
    K, invK = O.make_K(64, 32)
    d = dev()
    xg = x.to(d).requires_grad_(True)
    dg = disps[-1].to(d).requires_grad_(True)
    rg = [r.to(d).requires_grad_(True) for r in rv]
    tg = [t.to(d).requires_grad_(True) for t in tv]
    ch = torch.zeros(1, 2, 32, 64, 3, dtype=torch.int32, device=d)
    loss = M.view_synthesis_loss(xg, [dg], rg, tg, K.to(d), invK.to(d), scales=(1.0,), auto_loss=torch.zeros(2, 1, 32, 64, device=d),
                                 debug_choices=ch)
    loss.backward()
    assert ((ch[..., 0] & 3) == 0).all()                       # the automask won everywhere
    assert all(g.grad.abs().max().item() == 0.0 for g in rg + tg) and xg.grad.abs().max().item() == 0.0
    dd = disps[-1].double().requires_grad_(True)                 # what is left is the smoothness gradient
    nd = (dd / (dd.mean(dim=(2, 3), keepdim=True) + 1e-7))[:, 0]
    (O.smooth_loss(nd, x.double()[:, 1]) * 1e-3).backward()
    assert rel_max(dg.grad.cpu(), dd.grad) <= GRAD_RTOL
    # (b) identical sources, identical (non-inverted) poses
    x2 = x.clone(); x2[:, 2] = x2[:, 0]
    xg = x2.to(d).requires_grad_(True)
    rg = [rv[1].to(d).requires_grad_(True), rv[1].to(d).requires_grad_(True)]
    tg = [tv[1].to(d).requires_grad_(True), tv[1].to(d).requires_grad_(True)]
    ch.zero_()
    loss = M.view_synthesis_loss(xg, [disps[-1].to(d).requires_grad_(True)], rg, tg, K.to(d), invK.to(d), scales=(1.0,), invert=[False, False],
                                 debug_choices=ch)
    loss.backward()
    assert ((ch[..., 0] & 3) == 1).all()                       # source 0 everywhere
    assert rg[1].grad.abs().max().item() == 0.0 and tg[1].grad.abs().max().item() == 0.0 and xg.grad[:, 2].abs().max().item() == 0.0
    assert rg[0].grad.abs().max().item() > 0.0 and xg.grad[:, 0].abs().max().item() > 0.0


def test_two_by_two_images():
    """reflect padding needs W, H >= 2; the reference's own tests use 2x2 inputs (test/runtests.jl:52-68)"""
    torch.manual_seed(0)
    x = torch.rand(1, 3, 1, 2, 2)
    disps = [torch.rand(1, 1, 2, 2) * 0.5 + 0.2]
    rv = [torch.tensor([[0.0, 0.0, 0.01]]), torch.tensor([[0.01, 0.0, 0.0]])]
    tv = [torch.zeros(1, 3), torch.zeros(1, 3)]
    K, invK = O.make_K(2, 2, f=5.0)
    d = dev()
    dg = [t.to(d).requires_grad_(True) for t in disps]
    loss = M.view_synthesis_loss(x.to(d), dg, [r.to(d) for r in rv], [t.to(d) for t in tv], K.to(d), invK.to(d), scales=(1.0,))
    loss.backward()
    ref = oracle_vsl(x, disps, rv, tv, K, invK, scales=(1.0,))
    assert abs(loss.item() - ref["loss"]) < 1e-5 * abs(ref["loss"])
    assert rel_max(dg[0].grad.cpu(), ref["gdisp"][0]) <= 1e-3     # (4 pixels: one float32 tie moves everything)


class _StubModel:
    """stands in for the reference's Model (src/model.jl:8-20): returns fixed disparities (leaf tensors) and poses"""

    def __init__(self, disps, rv, tv):
        self.disps, self.poses = disps, [M.Pose(r, t) for r, t in zip(rv, tv)]

    def __call__(self, x, source_ids, target_id):
        assert list(source_ids) == [0, 2] and target_id == 1
        return self.disps, self.poses


@pytest.mark.parametrize("automasking", [False, True])
def test_train_loss_with_params_and_cache(automasking):
    """train_loss(model, x, auto_loss, cache, parameters, do_visualization) -> (loss, vis_disparity, vis_warped, vis_loss),
    src/training.jl:21-78, with the containers of src/Monodepth.jl:32-55"""
    N, C, H, W = 3, 3, 48, 96
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=8)
    K, invK = O.make_K(W, H)
    d = dev()
    params = M.Params(target_size=(W, H), batch_size=N, automasking=automasking)
    cache = M.TrainCache(M.SSIM(), M.Backproject(W, H), M.Project(W, H), K.to(d), invK.to(d), 1, (0, 2), (0.125, 0.25, 0.5, 1.0))
    xg = x.to(d)
    dg = [t.to(d).requires_grad_(True) for t in disps]
    rg = [t.to(d).requires_grad_(True) for t in rv]
    tg = [t.to(d).requires_grad_(True) for t in tv]
    auto = M.automasking_loss(cache.ssim, xg, xg[:, cache.target_id], cache.source_ids)
    loss, vis_disp, vis_warped, vis_loss = M.train_loss(_StubModel(dg, rg, tg), xg, auto, cache, params, True)
    loss.backward()
    xd = x.double()
    auto_ref = O.automasking_loss(O.SSIM(), xd, xd[:, 1], (0, 2))
    ref_loss, (ref_warped, ref_wl) = O.view_synthesis_loss(xd, [t.double() for t in disps], [t.double() for t in rv], [t.double() for t in tv],
                                                           K.double(), invK.double(), automasking=automasking, auto_loss=auto_ref, return_viz=True)
    assert abs(loss.item() - ref_loss.item()) <= LOSS_RTOL * abs(ref_loss.item())
    assert not vis_disp.is_cuda and torch.equal(vis_disp, disps[-1])                      # cpu(disparity) of the last scale
    assert len(vis_warped) == 2 and all(not w.is_cuda for w in vis_warped) and not vis_loss.is_cuda
    for a, b in zip(vis_warped, ref_warped):
        assert torch.allclose(a.double(), b, atol=2e-6)
    assert ((vis_loss.double() - ref_wl).abs() > 1e-5).double().mean().item() < 1e-3      # (arg-min / automask ties aside)
    assert all(t.grad is not None and torch.isfinite(t.grad).all() for t in dg + rg + tg)
    # without visualisation the three extra returns are None, and the loss is the same number
    loss2, a, b, c = M.train_loss(_StubModel(dg, rg, tg), xg, auto, cache, params, False)
    assert a is None and b is None and c is None and loss2.item() == loss.item()
