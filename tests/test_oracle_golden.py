"""The reference's own testsets (test/runtests.jl) restated against the oracle.

Every known-answer vector the reference holds for the hot path is checked here;
this is what pins oracle/torch_oracle.py.
"""
import math

import numpy as np
import torch

from oracle import torch_oracle as O

torch.manual_seed(0)
F64 = torch.float64


def _rodrigues(v):
    # independent rotation-vector -> matrix (scipy stands in for Rotations.RotationVec)
    from scipy.spatial.transform import Rotation
    return torch.tensor(Rotation.from_rotvec(np.asarray(v)).as_matrix(), dtype=F64)


def test_rotations():  # test/runtests.jl:14-29
    v = torch.rand(1, 3, dtype=F64)
    src = O.so3_exp_map(v)
    assert torch.allclose(src[0], _rodrigues(v[0].numpy()), atol=1e-5)


def test_hat_rrule():  # test/runtests.jl:21 (test_rrule(hat, v): finite differences)
    v = torch.rand(2, 3, dtype=F64, requires_grad=True)
    d = torch.rand(2, 3, 3, dtype=F64)
    (g,) = torch.autograd.grad((O.hat(v) * d).sum(), v)
    assert torch.allclose(g, O.hat_pullback(d), atol=1e-12)
    assert torch.autograd.gradcheck(O.hat, (v,))


def test_transformation():  # test/runtests.jl:31-50
    rvec = torch.rand(1, 3, dtype=F64)
    tvec = torch.rand(1, 3, dtype=F64)
    p = torch.rand(3, dtype=F64)
    R, t = O.composeT(rvec, tvec, False)
    tp = _rodrigues(rvec[0].numpy()) @ p + t[0]
    np_ = R[0] @ p + t[0]
    assert torch.allclose(np_, tp, atol=1e-6)
    R, t = O.composeT(rvec, tvec, True)
    invR = _rodrigues(rvec[0].numpy()).t()
    invt = -(invR @ tvec[0])
    assert torch.allclose(R[0] @ np_ + t[0], invR @ np_ + invt, atol=1e-6)
    assert torch.allclose(R[0] @ np_ + t[0], p, atol=1e-6)


def test_ssim():  # test/runtests.jl:52-68
    ssim = O.SSIM()
    one = torch.ones(1, 1, 2, 2, dtype=F64)
    assert torch.allclose(ssim(one, one), torch.zeros_like(one))
    score = ssim(one, torch.zeros_like(one))
    assert torch.allclose(score, torch.full_like(one, 0.5), atol=1e-1)
    assert torch.allclose(score, torch.full_like(one, 0.49995000499950004), atol=1e-14)
    a, b = torch.rand(2, 1, 2, 2, dtype=F64), torch.rand(2, 1, 2, 2, dtype=F64)
    assert torch.allclose(ssim(a, b), ssim(b, a))


def test_smooth_loss():  # test/runtests.jl:70-83
    # Julia: reshape(transpose(reshape(0:0.1:0.3,(2,2))),(2,2,1,1)) -> disp[w,h]: [0 .1; .2 .3]
    # i.e. disp[w=1,:]=(0,.1), disp[w=2,:]=(.2,.3)  => torch (H,W) = [[0,.2],[.1,.3]]
    disp = torch.tensor([[0.0, 0.2], [0.1, 0.3]], dtype=F64).reshape(1, 2, 2)
    image = torch.ones(1, 1, 2, 2, dtype=F64)
    sl = O.smooth_loss(disp, image)
    tl = (disp[:, :, :-1] - disp[:, :, 1:]).abs().mean() + (disp[:, :-1] - disp[:, 1:]).abs().mean()
    assert torch.allclose(sl, tl)
    assert abs(sl.item() - 0.3) < 1e-12
    image = torch.tensor([[0.1, 0.3], [0.2, 0.4]], dtype=F64).reshape(1, 1, 2, 2)
    sl = O.smooth_loss(disp, image)
    assert abs(sl.item() - 0.2542) < 1e-4
    assert abs(sl.item() - (0.2 * math.exp(-0.2) + 0.1 * math.exp(-0.1))) < 1e-12


def test_disparity_to_depth():  # test/runtests.jl:85-92
    disp = torch.rand(2, 32, 32, dtype=F64)
    depth = O.disparity_to_depth(disp, 0.1, 100.0)
    assert depth.min() >= 0.1 and depth.max() <= 100.0


def test_identity_warp():  # test/runtests.jl:94-122 (default :zeros padding)
    res, N = 16, 2
    torch.manual_seed(1)
    image = torch.rand(N, 1, res, res, dtype=F64)
    # the reference draws depth in [0,1); depths below ~1e-4 make its own test fail through the
    # 1e-7 in the perspective divide (src/utils.jl:97), so keep away from 0 here
    depth = torch.rand(N, res * res, dtype=F64) * 0.9 + 0.1
    K = torch.tensor([[910.0, 0, res / 2], [0, 910.0, res / 2], [0, 0, 1]], dtype=F64)
    invK = torch.linalg.inv(K)
    R = O.so3_exp_map(torch.zeros(N, 3, dtype=F64))
    t = torch.zeros(N, 3, dtype=F64)
    pts = O.Backproject(res, res)(depth, invK)
    uv = O.Project(res, res)(pts, K, R, t).reshape(N, res, res, 2)
    for mode in ("zeros", "border"):
        sampled = O.grid_sample(image, uv, padding_mode=mode)
        assert torch.allclose(image, sampled, atol=1e-3)


def test_pose_derivative():  # test/runtests.jl:124-142 (values from SURVEY.md section 4)
    x = torch.tensor([3.0, 2.0, 1.0], dtype=F64)
    target = torch.tensor([1.0, 2.0, 3.0], dtype=F64)
    r = torch.tensor([[1.0, 0.0, 0.0]], dtype=F64, requires_grad=True)
    t = torch.zeros(1, 3, dtype=F64, requires_grad=True)
    R = O.so3_exp_map(r)
    l = torch.sqrt((((R[0] @ x) + t[0] - target) ** 2).sum())
    gr, gt = torch.autograd.grad(l, (r, t))
    assert abs(l.item() - 2.775608012559207) < 1e-12
    assert torch.allclose(gr[0], torch.tensor([1.3435210063, 1.1003665905, -2.8688708364], dtype=F64), atol=1e-9)
    assert torch.allclose(gt[0], torch.tensor([0.7205628428, -0.6344074398, -0.2798506565], dtype=F64), atol=1e-9)


def test_min_tie_routing():  # first index wins ties
    a = torch.zeros(1, 1, 2, 2, dtype=F64, requires_grad=True)
    b = torch.zeros(1, 1, 2, 2, dtype=F64, requires_grad=True)
    O.apply_mask(a, b).sum().backward()
    assert torch.all(a.grad == 1) and torch.all(b.grad == 0)


def test_full_loss_runs_and_differentiates():
    x, disps, rv, tv = O.synthetic_batch(2, 3, 32, 64, dtype=F64)
    K, invK = O.make_K(64, 32, dtype=F64)
    for d in disps:
        d.requires_grad_(True)
    for r in rv + tv:
        r.requires_grad_(True)
    auto = O.automasking_loss(O.SSIM(), x, x[:, 1], (0, 2))
    loss = O.view_synthesis_loss(x, disps, rv, tv, K, invK, automasking=True, auto_loss=auto)
    loss.backward()
    assert torch.isfinite(loss) and all(torch.isfinite(d.grad).all() for d in disps)
    assert all(torch.isfinite(r.grad).all() for r in rv + tv)
