"""The C ABI used from plain C (tests/cabi/harness.c, gcc against include/md2.h, linked to libmd2_b200.so): create ->
device-pointer fwdbwd -> host-pointer fwdbwd -> error path -> destroy; results must equal the ctypes path bit for bit
(same library, same kernels) and the oracle within the stated bars."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

import monodepth2_jl_b200 as M
from oracle import torch_oracle as O
from util import check_vsl_statistical, oracle_vsl

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_harness(tmp):
    exe = os.path.join(tmp, "harness")
    libdir = os.path.dirname(M.LIB_PATH)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "cabi", "harness.c"), "-o", exe, "-L", libdir, "-lmd2_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_c_harness_matches_ctypes_and_oracle(tmp_path):
    M.load_library()
    exe = build_harness(str(tmp_path))
    N, C, H, W, L = 3, 3, 64, 128, 4
    x, disps, rv, tv = O.synthetic_batch(N, C, H, W, seed=21)
    K, invK = O.make_K(W, H)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("5i", N, C, H, W, L))
        for t in [x] + list(disps) + [rv[0], tv[0], rv[1], tv[1], K.t().contiguous(), invK.t().contiguous()]:
            f.write(t.contiguous().numpy().astype(np.float32).tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok launches=3 ")
    raw = np.fromfile(fout, dtype=np.float32)
    sizes = [1] + [d.numel() for d in disps] + [3 * N] * 4
    assert raw.size == 2 * sum(sizes)

    def block(k):
        out, off = [], k * sum(sizes)
        for n in sizes:
            out.append(torch.from_numpy(raw[off:off + n].copy())); off += n
        return out

    dev = torch.device("cuda", 0)
    dg = [d.to(dev).requires_grad_(True) for d in disps]
    rg = [t.to(dev).requires_grad_(True) for t in rv]
    tg = [t.to(dev).requires_grad_(True) for t in tv]
    loss = M.view_synthesis_loss(x.to(dev), dg, rg, tg, K.to(dev), invK.to(dev))
    loss.backward()
    mine = [loss.detach().reshape(1)] + [d.grad.reshape(-1) for d in dg] + [rg[0].grad.reshape(-1), tg[0].grad.reshape(-1), rg[1].grad.reshape(-1), tg[1].grad.reshape(-1)]
    b0, b1 = block(0), block(1)
    for a, b in zip(b0, mine):
        assert torch.equal(a, b.cpu())                         # device entry point from C == from ctypes, bit for bit
    ref = oracle_vsl(x, disps, rv, tv, K, invK)
    for blk in (b0, b1):
        out = dict(loss=float(blk[0]), gdisp=[g.reshape(d.shape) for g, d in zip(blk[1:1 + L], disps)],
                   grvec=[blk[1 + L].reshape(N, 3), blk[3 + L].reshape(N, 3)], gtvec=[blk[2 + L].reshape(N, 3), blk[4 + L].reshape(N, 3)])
        check_vsl_statistical(out, ref, tag="C harness")
