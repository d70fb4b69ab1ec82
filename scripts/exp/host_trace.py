"""experiment: stage time stamps of the host-buffer entry point (MD2_HOST_TRACE=1, eager first call)"""
import os, sys
os.environ["MD2_HOST_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import monodepth2_jl_b200 as M
from monodepth2_jl_b200 import synthetic as SY
NB, CH, H, W = 8, 1, 128, 416
x, d, r, t = SY.synthetic_batch(NB, CH, H, W, seed=42)
K, invK = SY.make_K(W, H)
for groups in (1, 2, 4):
    print("groups", groups, file=sys.stderr)
    hv = M.HostViewSynthesisLoss(NB, CH, H, W, [(q.shape[-1], q.shape[-2]) for q in d], K, invK, groups=groups)
    hv(x, d, r, t)
    # change the descriptor key so that the next call is eager again (new object = new pinned buffers)
