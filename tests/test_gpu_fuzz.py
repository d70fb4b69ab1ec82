"""Odd shapes through the fused calls (value + gradient, forward-only, automask formed inside) against the float64 oracle:
image sides that are multiples of nothing, every count of decoder scales, C = 1 / 3, with and without automasking."""
import random

import pytest
import torch

import monodepth2_jl_b200 as M
from oracle import torch_oracle as O
from util import check_vsl_statistical, oracle_vsl

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
    ref = oracle_vsl(x, disps, rv, tv, K, invK, automask=am, scales=scales)
    xg = x.to(dev).requires_grad_(True)
    dg = [d.to(dev).requires_grad_(True) for d in disps]
    rg = [r.to(dev).requires_grad_(True) for r in rv]
    tg = [t.to(dev).requires_grad_(True) for t in tv]
    loss = M.view_synthesis_loss(xg, dg, rg, tg, K.to(dev), invK.to(dev), scales=scales, compute_automask=am)
    loss.backward()
    out = dict(loss=loss.item(), gdisp=[d.grad for d in dg], grvec=[r.grad for r in rg], gtvec=[t.grad for t in tg], gx=xg.grad)
    check_vsl_statistical(out, ref, tag=f"{W}x{H}x{N} C={C} L={L} am={am}", frac=0.99, pose_rtol=2e-2)
    with torch.no_grad():       # forward-only kernel
        l2 = M.view_synthesis_loss(xg.detach(), [d.detach() for d in dg], [r.detach() for r in rg], [t.detach() for t in tg],
                                   K.to(dev), invK.to(dev), scales=scales, compute_automask=am)
    assert abs(l2.item() - ref["loss"]) <= 1e-5 * abs(ref["loss"])
