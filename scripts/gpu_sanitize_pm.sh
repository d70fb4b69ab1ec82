#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the copy-engine automask pre-pass (photomin_fwd_bulk_kernel: cp.async.bulk row
# copies on mbarriers, double-buffered passes), interior + border tiles, alone and inside the fused call
TAG=${1:-r2}
mkdir -p gpurun_out
cat > /tmp/san_pm.py <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import monodepth2_jl_b200 as M
from oracle import torch_oracle as O
dev = torch.device("cuda", 0)
for (N, C, H, W) in [(2, 3, 40, 104), (1, 3, 70, 160)]:
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=3)
    K, invK = O.make_K(W, H)
    xg = x.to(dev)
    auto = M.automasking_loss(M.SSIM(), xg, xg[:, 1], (0, 2))
    ref = O.automasking_loss(O.SSIM(), x.double(), x.double()[:, 1], (0, 2))
    err = (auto.cpu().double() - ref).abs().max().item()
    loss = M.view_synthesis_loss(xg, [d.to(dev) for d in disps], [r.to(dev) for r in rv], [t.to(dev) for t in tv], K.to(dev), invK.to(dev), compute_automask=True)
    torch.cuda.synchronize()
    print("ok", N, C, H, W, "max err of the map", err, "loss", float(loss))
    assert err < 1e-5
PY
for TOOL in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $TOOL python /tmp/san_pm.py > gpurun_out/${TAG}_sanitizer_pm_${TOOL}.log 2>&1; echo "$TOOL rc=$?"
  tail -3 gpurun_out/${TAG}_sanitizer_pm_${TOOL}.log
done
