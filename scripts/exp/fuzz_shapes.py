"""Random small shapes through the fused call against the float64 oracle (statistical bars of tests/util.py)."""
import os, sys, random, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import monodepth2_jl_b200 as M
from oracle import torch_oracle as O
from util import check_vsl_statistical, oracle_vsl
dev = torch.device("cuda", 0)
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    W, H, N, C = rng.randint(18, 150), rng.randint(10, 80), rng.randint(1, 3), rng.choice([1, 3])
    L = rng.randint(1, 4)
    scales = tuple([0.125, 0.25, 0.5, 1.0][4 - L:])
    am = rng.random() < 0.4
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, scales=scales, seed=1000 + trial)
    K, invK = O.make_K(W, H)
    try:
        ref = oracle_vsl(x, disps, rv, tv, K, invK, automask=am, scales=scales)
        xg = x.to(dev).requires_grad_(True)
        dg = [d.to(dev).requires_grad_(True) for d in disps]
        rg = [r.to(dev).requires_grad_(True) for r in rv]; tg = [t.to(dev).requires_grad_(True) for t in tv]
        loss = M.view_synthesis_loss(xg, dg, rg, tg, K.to(dev), invK.to(dev), scales=scales, compute_automask=am)
        loss.backward()
        out = dict(loss=loss.item(), gdisp=[d.grad for d in dg], grvec=[r.grad for r in rg], gtvec=[t.grad for t in tg], gx=xg.grad)
        check_vsl_statistical(out, ref, tag=f"{W}x{H}x{N} C={C} L={L} am={am}", frac=0.99, pose_rtol=2e-2)
        with torch.no_grad():
            l2 = M.view_synthesis_loss(xg.detach(), [d.detach() for d in dg], [r.detach() for r in rg], [t.detach() for t in tg], K.to(dev), invK.to(dev), scales=scales, compute_automask=am)
        assert abs(l2.item() - ref["loss"]) <= 1e-5 * abs(ref["loss"]), ("fwd-only", l2.item(), ref["loss"])
        print("ok  ", W, H, N, C, L, am, f"{loss.item():.6f}")
    except Exception as e:
        bad += 1
        print("FAIL", W, H, N, C, L, am, repr(e)[:300])
print("failures:", bad)
