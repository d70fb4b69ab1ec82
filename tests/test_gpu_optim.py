"""The optimiser side of the path on the GPU: the fused Flux-ADAM kernel (md2_adam_step) and the reference's triplet
optimiser `slow_depth` (src/simple_depth.jl:1-62, md2_slow_depth) against the CPU oracle loop."""
import pytest
import torch

import monodepth2_jl_b200 as M
from oracle import torch_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("shapes", [[(64, 8)], [(7,), (3, 5), (128,)], [(1 << 16,), (4,), (12,)]])
def test_adam_step_matches_flux_rule(shapes):
    """(vectorised path: every count a multiple of 4; scalar path: ragged counts; several steps so that the bias
    corrections b^t are exercised)"""
    torch.manual_seed(0)
    ps = [torch.randn(*s) for s in shapes]
    ref_p = [p.double().clone() for p in ps]
    ref = O.FluxAdam(ref_p, eta=3e-3)
    gp = [p.to(dev()).clone() for p in ps]
    opt = M.Adam(gp, lr=3e-3)
    for it in range(7):
        gs = [torch.randn(*s) * (10.0 ** (it - 3)) for s in shapes]
        ref.step([g.double() for g in gs])
        opt.step([g.to(dev()) for g in gs])
    assert opt.steps == 7
    for a, b in zip(gp, ref_p):
        assert torch.allclose(a.cpu().double(), b, rtol=2e-6, atol=1e-7)


def test_adam_grad_scale_and_checkpoint_round_trip():
    torch.manual_seed(1)
    p0 = torch.randn(1000)
    g = [torch.randn(1000) for _ in range(6)]
    a = p0.to(dev()).clone(); oa = M.Adam([a], lr=1e-2)
    for k in range(6):
        oa.step([g[k].to(dev()) * 4.0], grad_scale=0.25)      # (SUM over 4 ranks, then the mean)
    b = p0.to(dev()).clone(); ob = M.Adam([b], lr=1e-2)
    for k in range(3):
        ob.step([g[k].to(dev())])
    sd = ob.state_dict()                                       # checkpoint after 3 steps ...
    c = b.clone(); oc = M.Adam([c], lr=1.0)
    oc.load_state_dict(sd)                                     # ... resumed by a fresh optimiser
    for k in range(3, 6):
        oc.step([g[k].to(dev())])
    assert oc.steps == 6
    assert torch.allclose(a, c, rtol=1e-6, atol=1e-7)


def test_slow_depth_matches_the_oracle_loop():
    """config 1 in small: 30 iterations of the triplet optimiser from the reference's start point"""
    x, _, _, _ = O.synthetic_batch(1, 3, 48, 96, seed=3)
    K, invK = O.make_K(96, 48)
    iters = 30
    rd, rr, rt, rh = O.slow_depth(x.double(), K.double(), invK.double(), iters=iters)
    logged = []
    disp, poses, hist = M.slow_depth(x.to(dev()), K.to(dev()), invK.to(dev()), iters=iters, on_log=lambda i, d, p: logged.append(i))
    assert logged == [1, 5, 10, 15, 20, 25, 30]                                  # src/simple_depth.jl:23
    hist = hist.cpu().double()
    ref = torch.tensor(rh, dtype=torch.float64)
    assert abs(hist[0] - ref[0]) <= 1e-5 * ref[0]
    # ADAM's steps are sign-like while v is young (|step| ~ eta whatever |g| is): elements whose gradient is at rounding
    # level move the other way in float32 and the trajectories separate slowly -- the reference's own float32 run does
    # the same against float64, so the bar on the trajectory is looser than the bar on one evaluation
    assert torch.allclose(hist, ref, rtol=1e-3), (hist - ref).abs().max()
    assert ref[-1] < ref[0] and hist[-1] < hist[0]
    assert (disp.cpu().double() - rd).abs().max() <= 2 * 3e-4 * iters
    assert (disp.cpu().double() - rd).abs().mean() < 1.5 * 3e-4                # (a few flipped sign-like steps per element)
    for p, r, t in zip(poses, rr, rt):
        assert torch.allclose(p.rvec.cpu().double(), r, atol=1e-4) and torch.allclose(p.tvec.cpu().double(), t, atol=1e-4)


def test_slow_depth_is_the_loop_of_its_parts():
    """md2_slow_depth (graph replay on the device) == a host loop of { fused value + gradient; md2_adam_step }, bit for
    bit: the loss kernels are deterministic without the source-image scatter"""
    x, _, _, _ = O.synthetic_batch(1, 3, 32, 64, seed=7)
    K, invK = O.make_K(64, 32)
    d = dev()
    xg = x.to(d)
    iters = 10
    disp = torch.full((1, 1, 32, 64), 0.5, device=d).requires_grad_(True)
    rv = [torch.tensor([[0.0, 0.0, 0.01]], device=d).requires_grad_(True) for _ in range(2)]
    tv = [torch.zeros(1, 3, device=d).requires_grad_(True) for _ in range(2)]
    theta = [disp, rv[0], tv[0], rv[1], tv[1]]
    opt = M.Adam([t.detach() for t in theta], lr=3e-4)       # (same storage: detach() shares it)
    hist = []
    for _ in range(iters):
        for t in theta:
            t.grad = None
        loss = M.simple_depth_loss(xg, disp, [M.Pose(r, t) for r, t in zip(rv, tv)], K.to(d), invK.to(d))
        loss.backward()
        opt.step([t.grad.contiguous() for t in theta])
        hist.append(loss.detach())
    d2, p2, h2 = M.slow_depth(xg, K.to(d), invK.to(d), iters=iters)
    assert torch.equal(torch.stack(hist), h2)
    assert torch.equal(disp.detach(), d2) and torch.equal(rv[1].detach(), p2[1].rvec) and torch.equal(tv[0].detach(), p2[0].tvec)


def test_slow_depth_resumes_in_chunks():
    """one call of 12 iterations == 12 calls' worth through the logging path (same graph, same device clock)"""
    x, _, _, _ = O.synthetic_batch(1, 1, 32, 64, seed=5)
    K, invK = O.make_K(64, 32)
    d1, p1, h1 = M.slow_depth(x.to(dev()), K.to(dev()), invK.to(dev()), iters=12)
    d2, p2, h2 = M.slow_depth(x.to(dev()), K.to(dev()), invK.to(dev()), iters=12, log_step=2, on_log=lambda *a: None)
    assert torch.equal(h1, h2) and torch.equal(d1, d2) and torch.equal(p1[0].rvec, p2[0].rvec)


def test_slow_depth_at_config1_size_matches_the_oracle():
    """BASELINE.json configs[0] at its real size: a 416x128 RGB triplet, first iterations of the 500"""
    x, _, _, _ = O.synthetic_batch(1, 3, 128, 416, seed=42)
    K, invK = O.make_K(416, 128)
    iters = 6
    rd, rr, rt, rh = O.slow_depth(x.double(), K.double(), invK.double(), iters=iters)
    disp, poses, hist = M.slow_depth(x.to(dev()), K.to(dev()), invK.to(dev()), iters=iters)
    ref = torch.tensor(rh, dtype=torch.float64)
    assert abs(hist[0].item() - ref[0].item()) <= 1e-5 * ref[0].item()
    assert torch.allclose(hist.cpu().double(), ref, rtol=2e-4)
    assert (disp.cpu().double() - rd).abs().max() <= 2 * 3e-4 * iters
    for p, r, t in zip(poses, rr, rt):
        assert torch.allclose(p.rvec.cpu().double(), r, atol=5e-5) and torch.allclose(p.tvec.cpu().double(), t, atol=5e-5)
