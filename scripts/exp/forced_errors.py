"""diagnostic: forced-decision parity errors of the fused GPU kernel at one configuration"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import torch_oracle as O
from util import oracle_vsl_forced, rel_max
import test_gpu_forced as T
N, C, H, W, am = [int(v) for v in sys.argv[1:6]] if len(sys.argv) > 5 else (8, 1, 128, 416, 0)
x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=int(os.environ.get("SEED", "42")))
K, invK = O.make_K(W, H)
out = T.run_cuda(x, disps, rv, tv, K, invK, automask=bool(am))
ref = oracle_vsl_forced(x, disps, rv, tv, K, invK, out["choices"], auto=out["auto"].double() if am else None)
print("loss rel", abs(out["loss"] - ref["loss"]) / abs(ref["loss"]))
print("gdisp", [f"{rel_max(a, b):.2e}" for a, b in zip(out["gdisp"], ref["gdisp"])])
print("gx", f"{rel_max(out['gx'][:, [0, 2]], ref['gx'][:, [0, 2]]):.2e}")
print("grvec", [f"{rel_max(a, b):.2e}" for a, b in zip(out["grvec"], ref["grvec"])], "gtvec", [f"{rel_max(a, b):.2e}" for a, b in zip(out["gtvec"], ref["gtvec"])])
a, b = out["gdisp"][-1].double(), ref["gdisp"][-1]
err = (a - b).abs() / b.abs().max()
idx = torch.nonzero(err > 2e-4)
print("full-res outliers:", len(idx), idx[:12].tolist())
if len(idx):
    n, _, y, xx = idx[0].tolist()
    ch = out["choices"][-1, n]
    for dy in (-2, -1, 0, 1, 2):
        print([("%08x" % (ch[y + dy, xx + dx, 0].item() & 0xffffffff)) for dx in (-2, -1, 0, 1, 2) if 0 <= y + dy < H and 0 <= xx + dx < W])
    print("ours", a[n, 0, y - 1:y + 2, xx - 1:xx + 2], "ref", b[n, 0, y - 1:y + 2, xx - 1:xx + 2])
    w = ch[y, xx]
    print("word1", w[1].item() & 0x3fff, (w[1].item() >> 14) & 0x7fff, (w[1].item() >> 29) & 3, "word2", w[2].item() & 0x3fff, (w[2].item() >> 14) & 0x7fff, (w[2].item() >> 29) & 3)
if len(idx):
    # which single decision, flipped in the float64 oracle, explains the outlier?
    n, _, y, xx = idx[len(idx) // 2].tolist()
    base_err = err[n, 0, y, xx].item()
    xs, ds = x[n:n + 1], [d[n:n + 1] for d in disps]
    rvs, tvs = [r[n:n + 1] for r in rv], [t[n:n + 1] for t in tv]
    chn = out["choices"][:, n:n + 1].clone()
    scale_ref = b.abs().max()
    def err_with(chm):
        r = oracle_vsl_forced(xs, ds, rvs, tvs, K, invK, chm)
        return ((a[n, 0, y, xx] - r["gdisp"][-1][0, 0, y, xx] * 1.0 / N).abs() / scale_ref).item()
    print("per-image re-evaluation error:", err_with(chn), "batch:", base_err)
    L_ = chn.shape[0] - 1
    for dy in range(-2, 3):
        for dx in range(-2, 3):
            yy, xq = y + dy, xx + dx
            if not (0 <= yy < H and 0 <= xq < W):
                continue
            for what, word, mask in (("sel", 0, 3), ("pass0", 0, 4), ("pass1", 0, 8), ("l1s0", 0, 0x300), ("l1s1", 0, 0xc00), ("smx", 0, 0x300000), ("smy", 0, 0xc00000),
                                     ("mx0", 1, 1 << 29), ("my0", 1, 1 << 30), ("mx1", 2, 1 << 29), ("my1", 2, 1 << 30), ("x0_0", 1, 1), ("x0_1", 2, 1), ("y0_0", 1, 1 << 14), ("y0_1", 2, 1 << 14)):
                c2 = chn.clone()
                c2[L_, 0, yy, xq, word] ^= mask
                try:
                    e = err_with(c2)
                except RuntimeError:
                    continue
                if e < 0.3 * base_err:
                    print("flip", what, "at", (yy, xq), "-> error", e)
for i, (a_, b_) in enumerate(zip(out["gdisp"], ref["gdisp"])):
    e_ = (a_.double() - b_).abs() / b_.abs().max()
    j = torch.nonzero(e_ == e_.max())[0].tolist()
    big = torch.nonzero(e_ > 5e-5)
    print("scale", i, "shape", tuple(a_.shape), "max err at", j, "count>5e-5:", len(big), "rows:", sorted(set(big[:, 2].tolist()))[:12], "cols:", sorted(set(big[:, 3].tolist()))[:12],
          "ours", a_[tuple(j)].item(), "ref", b_[tuple(j)].item(), "max|ref|", b_.abs().max().item())
ga, gb = out["gx"][:, [0, 2]].double(), ref["gx"][:, [0, 2]]
ge = (ga - gb).abs() / gb.abs().max()
big = torch.nonzero(ge > 5e-5)
print("gx outliers > 5e-5:", len(big), big[:10].tolist(), "max|ref|", gb.abs().max().item())
for j in big[:4].tolist():
    print(j, "ours", ga[tuple(j)].item(), "ref", gb[tuple(j)].item())
