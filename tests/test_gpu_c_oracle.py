"""GPU parity against the SECOND oracle (oracle/c_oracle.c, plain-C float64 loops with a hand-derived reverse pass), through
the C ABI: the kernel's exported decisions are forced into the C oracle and EVERY gradient element has to be within
BASELINE.json's bars (loss 1e-5, gradients 1e-4), exactly as tests/test_gpu_forced.py does with the torch oracle -- and the
two float64 oracles have to agree with each other on those very decisions to 1e-9."""
import pytest
import torch

from oracle import c_oracle as CO
from oracle import torch_oracle as O
from test_gpu_forced import check_forced, run_cuda
from util import DECISION_FLIP_MAX, decision_mismatch, oracle_vsl_forced, rel_max

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,C,H,W,am", [(8, 1, 128, 416, False), (2, 3, 96, 160, True), (3, 1, 77, 61, True)])
def test_cuda_vs_c_oracle_forced(N, C, H, W, am):
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=43)
    K, invK = O.make_K(W, H)
    out = run_cuda(x, disps, rv, tv, K, invK, automask=am)
    auto = out["auto"].double() if am else None
    ref_c = CO.view_synthesis_loss(x, disps, rv, tv, K, invK, auto_loss=auto, choices=out["choices"])
    check_forced(out, ref_c, f"C oracle, forced {N},{C},{H},{W},{am}")
    ref_t = oracle_vsl_forced(x, disps, rv, tv, K, invK, out["choices"], auto=auto)
    assert abs(ref_c["loss"] - ref_t["loss"]) <= 1e-12 * abs(ref_t["loss"])
    for k in ("gdisp", "grvec", "gtvec"):
        for a, b in zip(ref_c[k], ref_t[k]):
            assert rel_max(a, b) <= 1e-9, (k, rel_max(a, b))
    assert rel_max(ref_c["gx"][:, [0, 2]], ref_t["gx"][:, [0, 2]]) <= 1e-9
    # the decisions themselves: the float64 oracle, left to decide on its own, takes the same branch at all but a handful of
    # pixels (forcing cannot hide a wrong gather cell / arg-min / mask: it would show here)
    own = CO.view_synthesis_loss(x, disps, rv, tv, K, invK, auto_loss=auto, grad=False, export_choices=True)
    mism = decision_mismatch(out["choices"], own["choices"], C)
    print("decision mismatch fractions", (N, C, H, W, am), mism)
    for kind, f in mism.items():
        assert f <= DECISION_FLIP_MAX, (kind, f)


def test_slow_depth_objective_vs_c_oracle():
    """config 1 (src/simple_depth.jl:25-41): value of the device objective against the C oracle"""
    import monodepth2_jl_b200 as M
    d = torch.device("cuda", 0)
    x, _, rv, tv = O.synthetic_batch(1, 3, 128, 416, seed=44)
    K, invK = O.make_K(416, 128)
    disp = torch.rand(1, 1, 128, 416) * 0.5 + 0.25
    ref = CO.simple_depth_loss(x, disp, rv, tv, K, invK, grad=False)
    got = M.simple_depth_loss(x.to(d), disp.to(d), [M.Pose(r.to(d), t.to(d)) for r, t in zip(rv, tv)], K.to(d), invK.to(d))
    assert abs(float(got) - ref["loss"]) <= 1e-5 * abs(ref["loss"]), (float(got), ref["loss"])
