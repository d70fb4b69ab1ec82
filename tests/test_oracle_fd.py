"""Finite-difference pinning of the oracle's derivatives: the float64 restatement is the
reference for every gradient parity test, so its autograd gradients are themselves checked against central differences of
its own value, along random directions, at well-conditioned points (margins to every discontinuity of the piecewise
smooth loss far larger than the step)."""
import torch

from oracle import torch_oracle as O
from util import conditioning

F64 = torch.float64


def _well_conditioned(N, C, H, W, automask, full_res):
    for seed in range(200):
        x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=seed, dtype=F64, full_res_disp=full_res)
        K, invK = O.make_K(W, H, dtype=F64)
        if conditioning(x, disps, rv, tv, K, invK, automask) > 2.0:       # margins of >= 2 float32 radii (~2e-5 px, ~1e-5 in the errors) >> the FD step
            return x, disps, rv, tv, K, invK
    raise RuntimeError("no well-conditioned seed")


def _directional_check(f, leaves, h=1e-7, tries=3, rtol=2e-5):
    val = f()
    grads = torch.autograd.grad(val, leaves)
    g = torch.Generator().manual_seed(0)
    for _ in range(tries):
        dirs = [torch.randn(t.shape, generator=g, dtype=F64) * t.detach().abs().mean().clamp_min(1e-3) for t in leaves]
        analytic = sum((a * d).sum() for a, d in zip(grads, dirs)).item()
        with torch.no_grad():
            for t, d in zip(leaves, dirs):
                t.add_(h * d)
            fp = f().item()
            for t, d in zip(leaves, dirs):
                t.sub_(2 * h * d)
            fm = f().item()
            for t, d in zip(leaves, dirs):
                t.add_(h * d)
        numeric = (fp - fm) / (2 * h)
        assert abs(numeric - analytic) <= rtol * max(abs(analytic), 1e-12), (numeric, analytic)


def test_train_loss_tail_gradients_match_finite_differences():
    for automask in (False, True):
        x, disps, rv, tv, K, invK = _well_conditioned(2, 1, 16, 24, automask, False)
        auto = O.automasking_loss(O.SSIM(), x, x[:, 1], (0, 2)) if automask else None
        leaves = [d.clone().requires_grad_(True) for d in disps] + [r.clone().requires_grad_(True) for r in rv] + \
                 [t.clone().requires_grad_(True) for t in tv]
        xr = x.clone().requires_grad_(True)
        L = len(disps)
        f = lambda: O.view_synthesis_loss(xr, leaves[:L], leaves[L:L + 2], leaves[L + 2:], K, invK, automasking=automask, auto_loss=auto)
        # every group of arguments on its own, then all together (incl. the source images)
        _directional_check(f, leaves[:L])
        _directional_check(f, leaves[L:])
        _directional_check(f, leaves + [xr])


def test_simple_depth_objective_gradients_match_finite_differences():
    x, disps, rv, tv, K, invK = _well_conditioned(1, 3, 16, 24, False, True)
    disp = disps[-1].clone().requires_grad_(True)
    r = [t.clone().requires_grad_(True) for t in rv]
    t_ = [t.clone().requires_grad_(True) for t in tv]
    f = lambda: O.simple_depth_loss(x, disp, r, t_, K, invK)
    _directional_check(f, [disp] + r + t_)


def test_operator_gradients_match_finite_differences():
    torch.manual_seed(0)
    a = (torch.rand(1, 2, 9, 11, dtype=F64) * 0.8 + 0.1).requires_grad_(True)
    b = (torch.rand(1, 2, 9, 11, dtype=F64) * 0.8 + 0.1).requires_grad_(True)
    w = torch.rand(1, 2, 9, 11, dtype=F64)
    _directional_check(lambda: (O.SSIM()(a, b) * w).sum(), [a, b])
    d = (torch.rand(2, 9, 11, dtype=F64) * 0.5 + 0.2).requires_grad_(True)
    img = torch.rand(2, 3, 9, 11, dtype=F64).requires_grad_(True)
    _directional_check(lambda: O.smooth_loss(d, img), [d, img], rtol=1e-4)
    rvec = (torch.randn(3, 3, dtype=F64) * 0.3).requires_grad_(True)
    tvec = torch.randn(3, 3, dtype=F64).requires_grad_(True)
    wR, wt = torch.randn(3, 3, 3, dtype=F64), torch.randn(3, 3, dtype=F64)
    for inv in (False, True):
        _directional_check(lambda: sum((o * ww).sum() for o, ww in zip(O.composeT(rvec, tvec, inv), (wR, wt))), [rvec, tvec])
