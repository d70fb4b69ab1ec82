#!/bin/bash
# compute-sanitizer (memcheck + racecheck + initcheck-free) over one small fused call per kernel family
TAG=${1:-r2}
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import monodepth2_jl_b200 as M
from oracle import torch_oracle as O
dev = torch.device("cuda", 0)
for (N, C, H, W, am) in [(2, 1, 40, 72, False), (2, 3, 33, 70, True)]:
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=3)
    K, invK = O.make_K(W, H)
    xg = x.to(dev).requires_grad_(True)
    dg = [d.to(dev).requires_grad_(True) for d in disps]
    rg = [r.to(dev).requires_grad_(True) for r in rv]; tg = [t.to(dev).requires_grad_(True) for t in tv]
    auto = M.automasking_loss(M.SSIM(), xg.detach(), xg.detach()[:, 1], (0, 2)) if am else None
    loss = M.view_synthesis_loss(xg, dg, rg, tg, K.to(dev), invK.to(dev), auto_loss=auto)
    loss.backward()
    if am:   # the automask map formed inside the call
        M.view_synthesis_loss(xg, dg, rg, tg, K.to(dev), invK.to(dev), compute_automask=True).backward()
    with torch.no_grad():
        l2 = M.view_synthesis_loss(xg.detach(), [d.detach() for d in dg], [r.detach() for r in rg], [t.detach() for t in tg], K.to(dev), invK.to(dev), auto_loss=auto)
    # stand-alone operators with tiles / merged scatter
    a, b = x[:, 0].to(dev).requires_grad_(True), x[:, 1].to(dev).requires_grad_(True)
    M.SSIM()(a, b).sum().backward()
    M.prediction_loss(M.SSIM(), [a, x[:, 2].to(dev).requires_grad_(True)], b).sum().backward()
    torch.cuda.synchronize()
    print("ok", N, C, H, W, am, float(loss.detach()), float(l2))
d, p, h = M.slow_depth(x[:1].to(dev), K.to(dev), invK.to(dev), iters=3)
torch.cuda.synchronize(); print("slow_depth ok", h.tolist())
PY
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL python /tmp/san.py > gpurun_out/${TAG}_sanitizer_${TOOL}.log 2>&1; echo "$TOOL rc=$?"
  tail -4 gpurun_out/${TAG}_sanitizer_${TOOL}.log
done
