import sys, os, ctypes as C, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import monodepth2_jl_b200 as M
from monodepth2_jl_b200 import synthetic as SY
dev = torch.device("cuda", 0)
W, H, N, Cc = 416, 128, 8, 1
x, disps, rv, tv = SY.synthetic_batch(N, Cc, H, W, seed=7)
K, invK = SY.make_K(W, H)
args = (x.to(dev), [d.to(dev) for d in disps], [r.to(dev) for r in rv], [t.to(dev) for t in tv], K.to(dev), invK.to(dev))
ctx = M.Context.get(dev)
for grad in (False, True):
    a = list(args)
    if grad:
        a[1] = [d.clone().requires_grad_(True) for d in a[1]]
    for _ in range(20):
        M.view_synthesis_loss(*a)
    ctx.profile(2)
    for _ in range(200):
        M.view_synthesis_loss(*a)
    p, m, f, n = ctx.profile_read_phases()
    ctx.profile(False)
    print("fwd+bwd" if grad else "fwd-only", "prep %.2f march %.2f finish %.2f us" % (1e3 * p / n, 1e3 * m / n, 1e3 * f / n))
