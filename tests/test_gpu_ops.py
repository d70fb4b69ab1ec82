"""GPU parity of every stand-alone operator (forward and backward, through the C ABI) against
the CPU oracle, including the reference's own testsets (test/runtests.jl) restated on the
CUDA operators."""
import math

import numpy as np
import pytest
import torch

import monodepth2_jl_b200 as M
from oracle import torch_oracle as O
from util import rel_l2, rel_max

pytestmark = pytest.mark.gpu
F64 = torch.float64


def dev():
    return torch.device("cuda", 0)


def pair(t):
    """(cuda float32 leaf, cpu float64 leaf) from a cpu float32 tensor"""
    return t.to(dev()).requires_grad_(True), t.double().requires_grad_(True)


def close(a, b, tol=1e-5):
    assert rel_max(a, b) <= tol, rel_max(a, b)


# ---------------- reference testsets on the CUDA operators ----------------
def test_ref_rotations_and_hat():
    from scipy.spatial.transform import Rotation
    v = torch.rand(4, 3)
    R = M.so3_exp_map(v.to(dev())).cpu()
    for n in range(4):
        assert np.allclose(R[n].numpy(), Rotation.from_rotvec(v[n].numpy()).as_matrix(), atol=1e-5)
    vg, vc = pair(v)
    d = torch.rand(4, 3, 3)
    (M.hat(vg) * d.to(dev())).sum().backward()
    assert torch.allclose(vg.grad.cpu().double(), O.hat_pullback(d.double()), atol=1e-6)
    assert torch.allclose(M.hat(vg).cpu().double(), O.hat(v.double()), atol=0)


def test_ref_transformation():
    rvec, tvec, p = torch.rand(3, 3), torch.rand(3, 3), torch.rand(3, 3)
    d = dev()
    R, t = M.composeT(rvec.to(d), tvec.to(d), False)
    Ro, to = O.composeT(rvec.double(), tvec.double(), False)
    assert torch.allclose(R.cpu().double(), Ro, atol=1e-6) and torch.allclose(t.cpu().double(), to, atol=1e-6)
    np_ = torch.einsum("nij,nj->ni", R.cpu(), p) + t.cpu()
    Ri, ti = M.composeT(rvec.to(d), tvec.to(d), True)
    back = torch.einsum("nij,nj->ni", Ri.cpu(), np_) + ti.cpu()
    assert torch.allclose(back, p, atol=1e-5)


def test_ref_ssim():
    d = dev()
    ssim = M.SSIM()
    one = torch.ones(1, 1, 2, 2, device=d)
    assert torch.allclose(ssim(one, one), torch.zeros_like(one), atol=1e-7)
    assert torch.allclose(ssim(one, torch.zeros_like(one)), torch.full_like(one, 0.49995000499950004), atol=1e-6)
    a, b = torch.rand(2, 1, 2, 2, device=d), torch.rand(2, 1, 2, 2, device=d)
    assert torch.allclose(ssim(a, b), ssim(b, a), atol=1e-6)


def test_ref_smooth_loss():
    d = dev()
    disp = torch.tensor([[0.0, 0.2], [0.1, 0.3]]).reshape(1, 2, 2).to(d)
    sl = M.smooth_loss(disp, torch.ones(1, 1, 2, 2, device=d))
    assert abs(sl.item() - 0.3) < 1e-6
    image = torch.tensor([[0.1, 0.3], [0.2, 0.4]]).reshape(1, 1, 2, 2).to(d)
    sl = M.smooth_loss(disp, image)
    assert abs(sl.item() - 0.2542) < 1e-4
    assert abs(sl.item() - (0.2 * math.exp(-0.2) + 0.1 * math.exp(-0.1))) < 1e-6


def test_ref_disparity_to_depth_range():
    depth = M.disparity_to_depth(torch.rand(2, 32, 32, device=dev()), 0.1, 100.0)
    assert depth.min().item() >= 0.1 - 1e-6 and depth.max().item() <= 100.0 + 1e-3


def test_ref_identity_warp():
    res, N = 16, 2
    d = dev()
    torch.manual_seed(1)
    image = torch.rand(N, 1, res, res)
    depth = torch.rand(N, res * res) * 0.9 + 0.1   # see tests/test_oracle_golden.py
    K = torch.tensor([[910.0, 0, res / 2], [0, 910.0, res / 2], [0, 0, 1]])
    invK = torch.linalg.inv(K.double()).float()
    R = M.so3_exp_map(torch.zeros(N, 3, device=d))
    t = torch.zeros(N, 3, device=d)
    pts = M.Backproject(res, res)(depth.to(d), invK.to(d))
    uv = M.Project(res, res)(pts, K.to(d).reshape(1, 3, 3), R, t).reshape(N, res, res, 2)
    for mode in ("zeros", "border"):
        sampled = M.grid_sample(image.to(d), uv, padding_mode=mode)
        assert torch.allclose(image, sampled.cpu(), atol=1e-3)


def test_ref_pose_derivative():
    x = torch.tensor([3.0, 2.0, 1.0], device=dev())
    target = torch.tensor([1.0, 2.0, 3.0], device=dev())
    r = torch.tensor([[1.0, 0.0, 0.0]], device=dev(), requires_grad=True)
    t = torch.zeros(1, 3, device=dev(), requires_grad=True)
    R = M.so3_exp_map(r)
    l = torch.sqrt((((R[0] @ x) + t[0] - target) ** 2).sum())
    l.backward()
    assert abs(l.item() - 2.775608012559207) < 1e-5
    assert torch.allclose(r.grad[0].cpu(), torch.tensor([1.3435210063, 1.1003665905, -2.8688708364]), atol=1e-5)
    assert torch.allclose(t.grad[0].cpu(), torch.tensor([0.7205628428, -0.6344074398, -0.2798506565]), atol=1e-5)


# ---------------- operator-by-operator forward + backward parity ----------------
def test_disparity_to_depth_grad():
    dg, dc = pair(torch.rand(2, 1, 8, 12))
    w = torch.rand(2, 1, 8, 12)
    (M.disparity_to_depth(dg, 0.1, 100.0) * w.to(dev())).sum().backward()
    (O.disparity_to_depth(dc, 0.1, 100.0) * w.double()).sum().backward()
    close(dg.grad, dc.grad)


def test_backproject_project_grads():
    W, H, N = 12, 8, 2
    K, invK = O.make_K(W, H)
    depth_g, depth_c = pair(torch.rand(N, W * H) * 5 + 1)
    rv = 0.05 * torch.randn(N, 3)
    tvv = 0.05 * torch.randn(N, 3)
    rg, rc = pair(rv)
    tg, tc = pair(tvv)
    w = torch.randn(N, W * H, 2)
    Rg, tug = M.composeT(rg, tg, True)
    uvg = M.Project(W, H)(M.Backproject(W, H)(depth_g, invK.to(dev())), K.to(dev()), Rg, tug)
    (uvg * w.to(dev())).sum().backward()
    Rc, tuc = O.composeT(rc, tc, True)
    uvc = O.Project(W, H)(O.Backproject(W, H)(depth_c, invK.double()), K.double(), Rc, tuc)
    (uvc * w.double()).sum().backward()
    close(uvg, uvc, 1e-5)
    close(depth_g.grad, depth_c.grad, 1e-4)
    close(rg.grad, rc.grad, 1e-4)
    close(tg.grad, tc.grad, 1e-4)


def test_so3_grad():
    rg, rc = pair(torch.randn(5, 3) * 0.5)
    w = torch.randn(5, 3, 3)
    (M.so3_exp_map(rg) * w.to(dev())).sum().backward()
    (O.so3_exp_map(rc) * w.double()).sum().backward()
    close(rg.grad, rc.grad, 1e-5)


@pytest.mark.parametrize("mode", ["zeros", "border"])
def test_grid_sample_grads(mode):
    N, Cc, H, W = 2, 3, 9, 13
    ig, ic = pair(torch.rand(N, Cc, H, W))
    grid = (torch.rand(N, 7, 11, 2) * 2.6 - 1.3)   # includes out-of-range coordinates
    gg, gc = pair(grid)
    w = torch.randn(N, Cc, 7, 11)
    og = M.grid_sample(ig, gg, padding_mode=mode)
    (og * w.to(dev())).sum().backward()
    oc = O.grid_sample(ic, gc, padding_mode=mode)
    (oc * w.double()).sum().backward()
    close(og, oc, 1e-5)
    close(ig.grad, ic.grad, 1e-5)
    close(gg.grad, gc.grad, 1e-4)


def test_upsample_grads():
    xg, xc = pair(torch.rand(2, 1, 6, 13))
    w = torch.randn(2, 1, 48, 104)
    og = M.upsample_bilinear(xg, (104, 48))
    (og * w.to(dev())).sum().backward()
    oc = O.upsample_bilinear(xc, (104, 48))
    (oc * w.double()).sum().backward()
    close(og, oc, 1e-5)
    close(xg.grad, xc.grad, 1e-5)


@pytest.mark.parametrize("shape", [(2, 3, 11, 17), (1, 1, 2, 2), (1, 1, 3, 3), (2, 1, 40, 70)])
def test_ssim_grads(shape):
    torch.manual_seed(1)
    xg, xc = pair(torch.rand(*shape))
    yg, yc = pair(torch.rand(*shape))
    w = torch.randn(*shape)
    og = M.SSIM()(xg, yg)
    (og * w.to(dev())).sum().backward()
    oc = O.SSIM()(xc, yc)
    (oc * w.double()).sum().backward()
    close(og, oc, 1e-5)
    close(xg.grad, xc.grad, 1e-4)
    close(yg.grad, yc.grad, 1e-4)


@pytest.mark.parametrize("C", [1, 3])
def test_photometric_and_prediction_loss_grads(C):
    torch.manual_seed(2)
    shape = (2, C, 14, 19)
    tg, tc = pair(torch.rand(*shape))
    p0g, p0c = pair(torch.rand(*shape))
    p1g, p1c = pair(torch.rand(*shape))
    w = torch.rand(2, 1, 14, 19)
    og = M.photometric_loss(M.SSIM(), p0g, tg)
    oc = O.photometric_loss(O.SSIM(), p0c, tc)
    close(og, oc, 1e-5)
    og = M.prediction_loss(M.SSIM(), [p0g, p1g], tg)
    (og * w.to(dev())).sum().backward()
    oc = O.prediction_loss(O.SSIM(), [p0c, p1c], tc)
    (oc * w.double()).sum().backward()
    close(og, oc, 1e-5)
    for a, b in ((p0g, p0c), (p1g, p1c), (tg, tc)):
        close(a.grad, b.grad, 1e-4)


def test_automasking_and_apply_mask():
    x = torch.rand(2, 3, 3, 12, 20)
    og = M.automasking_loss(M.SSIM(), x.to(dev()), x[:, 1].to(dev()), (0, 2))
    oc = O.automasking_loss(O.SSIM(), x.double(), x[:, 1].double(), (0, 2))
    close(og, oc, 1e-5)
    a = torch.zeros(1, 1, 2, 2, device=dev(), requires_grad=True)
    b = torch.zeros(1, 1, 2, 2, device=dev(), requires_grad=True)
    M._apply_mask(a, b).sum().backward()
    assert torch.all(a.grad == 1) and torch.all(b.grad == 0)   # mask wins ties


@pytest.mark.parametrize("C,H,W,off", [(3, 70, 160, 0), (1, 33, 100, 0), (3, 40, 99, 0), (3, 64, 128, 1), (1, 128, 416, 0), (3, 5, 72, 0)])
def test_automask_prepass_copy_engine_and_border_tiles(C, H, W, off):
    """photometric_min forward (the automask pre-pass): interior tiles are staged by cp.async.bulk row copies on mbarriers,
    tiles on the left / right border, widths that are not a multiple of 4 and frames that are not 16-byte aligned (off = 1:
    the batch starts one float into its allocation) through the threads -- every route against the oracle."""
    torch.manual_seed(5)
    N = 3
    buf = torch.rand(N * 3 * C * H * W + off)
    x = buf[off:].view(N, 3, C, H, W)
    xg = buf.to(dev())[off:].view(N, 3, C, H, W)
    assert xg.data_ptr() % 16 == (4 * off) % 16
    og = M.automasking_loss(M.SSIM(), xg, xg[:, 1], (0, 2))
    oc = O.automasking_loss(O.SSIM(), x.double(), x[:, 1].double(), (0, 2))
    close(og, oc, 1e-5)
    # one source (a single (channel, source) pass per channel: the other buffer parity pattern)
    og1 = M.automasking_loss(M.SSIM(), xg, xg[:, 1], (2,))
    oc1 = O.automasking_loss(O.SSIM(), x.double(), x[:, 1].double(), (2,))
    close(og1, oc1, 1e-5)


@pytest.mark.parametrize("normalize", [False, True])
def test_smooth_loss_grads(normalize):
    torch.manual_seed(3)
    dg, dc = pair(torch.rand(2, 10, 15))
    ig, ic = pair(torch.rand(2, 3, 10, 15))
    og = M.smooth_loss(dg, ig, normalize=normalize)
    (og * 1.7).backward()
    dn = dc / (dc.mean(dim=(1, 2), keepdim=True) + 1e-7) if normalize else dc
    oc = O.smooth_loss(dn, ic)
    (oc * 1.7).backward()
    assert abs(og.item() - oc.item()) < 1e-5 * abs(oc.item())
    close(dg.grad, dc.grad, 1e-4)
    close(ig.grad, ic.grad, 1e-4)


def test_warp_matches_oracle_and_grads():
    N, Cc, H, W = 2, 3, 24, 40
    x, disps, rv, tv = O.synthetic_batch(N, Cc, H, W, seed=4, full_res_disp=True)
    K, invK = O.make_K(W, H)
    d = dev()
    dg, dc = pair(disps[-1])
    rg = [pair(r) for r in rv]
    tg = [pair(t) for t in tv]
    xg, xc = pair(x)
    wts = [torch.randn(N, Cc, H, W) for _ in range(2)]
    Ps_g = [M.composeT(r[0], t[0], inv) for r, t, inv in zip(rg, tg, (True, False))]
    outs_g = M.warp(dg, xg, Ps_g, M.Backproject(W, H), M.Project(W, H), invK.to(d), K.to(d),
                    min_depth=0.1, max_depth=100.0, source_ids=(0, 2))
    sum(((o * w.to(d)).sum() for o, w in zip(outs_g, wts))).backward()
    Ps_c = [O.composeT(r[1], t[1], inv) for r, t, inv in zip(rg, tg, (True, False))]
    outs_c = O.warp(dc, xc, Ps_c, O.Backproject(W, H), O.Project(W, H), invK.double(), K.double(), 0.1, 100.0, (0, 2))
    sum(((o * w.double()).sum() for o, w in zip(outs_c, wts))).backward()
    for a, b in zip(outs_g, outs_c):
        close(a, b, 2e-5)
    err = ((dg.grad.cpu().double() - dc.grad).abs() / dc.grad.abs().max())
    assert (err <= 1e-4).double().mean().item() >= 0.995
    close(xg.grad, xc.grad, 2e-4)
    for (a, b) in rg + tg:
        close(a.grad, b.grad, 2e-3)
