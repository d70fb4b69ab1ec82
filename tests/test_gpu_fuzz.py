"""Odd shapes through the fused calls against the float64 oracle, strictly: the value + gradient call with the kernel's
discrete decisions forced onto the oracle (every gradient element within 1e-4, loss within 1e-5, tests/test_gpu_forced.py),
the forward-only call and the automask-formed-inside route on the same inputs.  Image sides that are multiples of nothing,
every count of decoder scales, C = 1 / 3, with and without automasking."""
import random

import pytest
import torch

import monodepth2_jl_b200 as M
from oracle import torch_oracle as O
from test_gpu_forced import check_forced
from util import LOSS_RTOL, oracle_vsl_forced

pytestmark = pytest.mark.gpu


def _cases(n, seed=0):
    rng = random.Random(seed)
    out = []
    for trial in range(n):
        W, H, N, C = rng.randint(18, 150), rng.randint(10, 80), rng.randint(1, 3), rng.choice([1, 3])
        L = rng.randint(1, 4)
        out.append((W, H, N, C, L, rng.random() < 0.4, 1000 + trial))
    return out


@pytest.mark.parametrize("W,H,N,C,L,am,seed", _cases(12))
def test_random_shapes_match_the_oracle(W, H, N, C, L, am, seed):
    dev = torch.device("cuda", 0)
    scales = tuple([0.125, 0.25, 0.5, 1.0][4 - L:])
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, scales=scales, seed=seed)
    K, invK = O.make_K(W, H)
    xg = x.to(dev).requires_grad_(True)
    dg = [d.to(dev).requires_grad_(True) for d in disps]
    rg = [r.to(dev).requires_grad_(True) for r in rv]
    tg = [t.to(dev).requires_grad_(True) for t in tv]
    auto = M.automasking_loss(M.SSIM(), xg.detach(), xg.detach()[:, 1], (0, 2)) if am else None
    ch = torch.zeros(L, N, H, W, 3, dtype=torch.int32, device=dev)
    loss = M.view_synthesis_loss(xg, dg, rg, tg, K.to(dev), invK.to(dev), scales=scales, auto_loss=auto, debug_choices=ch)
    loss.backward()
    out = dict(loss=loss.item(), gdisp=[d.grad.cpu() for d in dg], grvec=[r.grad.cpu() for r in rg], gtvec=[t.grad.cpu() for t in tg], gx=xg.grad.cpu())
    if am:
        assert torch.allclose(auto.cpu().double(), O.automasking_loss(O.SSIM(), x.double(), x.double()[:, 1], (0, 2)), atol=2e-6)
    ref = oracle_vsl_forced(x, disps, rv, tv, K, invK, ch.cpu(), auto=None if auto is None else auto.cpu().double(), scales=scales)
    check_forced(out, ref, f"{W}x{H}x{N} C={C} L={L} am={am}")
    with torch.no_grad():       # forward-only kernel, automask formed inside the call
        l2 = M.view_synthesis_loss(xg.detach(), [d.detach() for d in dg], [r.detach() for r in rg], [t.detach() for t in tg],
                                   K.to(dev), invK.to(dev), scales=scales, compute_automask=am)
    assert abs(l2.item() - loss.item()) <= LOSS_RTOL * abs(loss.item())
