import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import monodepth2_jl_b200 as M
from monodepth2_jl_b200 import synthetic as SY
from oracle import torch_oracle as O
dev = torch.device("cuda", 0)
W, H, N, C = 640, 192, 2, 3
x, disps, rv, tv = SY.synthetic_batch(N, C, H, W, seed=7)
xs = x.to(dev)
auto = M.automasking_loss(M.SSIM(), xs, xs[:, 1], (0, 2))
ref = O.automasking_loss(O.SSIM(), x.double(), x.double()[:, 1], (0, 2))
d = (auto.cpu().double() - ref).abs()
print("automask map: max abs diff", d.max().item(), "mean", d.mean().item(), "auto mean", auto.mean().item(), "ref mean", ref.mean().item())
K, invK = SY.make_K(W, H)
ch = torch.zeros(4, N, H, W, 3, dtype=torch.int32, device=dev)
dg = [t.to(dev).requires_grad_(True) for t in disps]
loss = M.view_synthesis_loss(xs, dg, [r.to(dev).requires_grad_(True) for r in rv], [t.to(dev).requires_grad_(True) for t in tv], K.to(dev), invK.to(dev), auto_loss=auto, debug_choices=ch)
sel = ch[..., 0] & 3
print("fraction of pixels where the automask wins, per scale:", [(sel[l] == 0).float().mean().item() for l in range(4)])
