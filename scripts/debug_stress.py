"""debug helper: stress-pose case, march kernel at several chunk heights vs tile kernel vs oracle"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from oracle import torch_oracle as O
    from util import oracle_vsl, rel_l2, rel_max
    from test_gpu_fused import run_cuda
    x, disps, rv, tv = O.synthetic_batch(2, 3, 48, 96, seed=5, pose_sigma=0.1)
    K, invK = O.make_K(96, 48)
    ref = oracle_vsl(x, disps, rv, tv, K, invK, automask=True)
    out = run_cuda(x, disps, rv, tv, K, invK, automask=True)
    print(os.environ.get("MD2_KERNEL"), os.environ.get("MD2_MARCH_ROWS"), "loss", out["loss"], ref["loss"])
    for nme in ("grvec", "gtvec"):
        print(" ", nme, [rel_max(a, b) for a, b in zip(out[nme], ref[nme])])
    print("  gdisp", [rel_l2(a, b) for a, b in zip(out["gdisp"], ref["gdisp"])], "gx", rel_l2(out["gx"][:, [0, 2]], ref["gx"][:, [0, 2]]))
    print("  grvec1", out["grvec"][1].tolist(), ref["grvec"][1].tolist())
    torch.save({k: out[k] for k in ("gdisp", "grvec", "gtvec", "gx")}, f"/tmp/out_{os.environ.get('MD2_KERNEL','march')}_{os.environ.get('MD2_MARCH_ROWS','0')}.pt")
else:
    for env in ({"MD2_KERNEL": "tile"}, {}, {"MD2_MARCH_ROWS": "8"}, {"MD2_MARCH_ROWS": "16"}, {"MD2_MARCH_ROWS": "48"}):
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, "child"], env=e)
    import torch
    a = torch.load("/tmp/out_tile_0.pt"); b = torch.load("/tmp/out_march_0.pt")
    for k in ("gdisp", "grvec", "gtvec"):
        for i, (u, v) in enumerate(zip(a[k], b[k])):
            d = (u - v).abs()
            print(k, i, "max diff", d.max().item(), "ref max", u.abs().max().item(), "n>1e-4", (d > 1e-4 * u.abs().max()).sum().item())
    d = (a["gx"] - b["gx"]).abs(); m = a["gx"].abs().max()
    idx = (d > 1e-4 * m).nonzero()
    print("gx diff count", len(idx), idx[:30].tolist())
    d = (a["gdisp"][3] - b["gdisp"][3]).abs(); m = a["gdisp"][3].abs().max()
    idx = (d > 1e-4 * m).nonzero()
    print("gdisp3 diff count", len(idx), idx[:30].tolist())
