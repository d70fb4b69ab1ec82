"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/md2.h declares, the ctypes mirror of md2_vsl_desc has the C layout, and the product
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest
import torch

import __graft_entry__ as G
import monodepth2_jl_b200 as M
from monodepth2_jl_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "md2.h")


@pytest.fixture(scope="module")
def lib():
    G.build()
    return M.load_library()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(md2_[a-zA-Z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/md2.h but not exported"
    assert set(L.EXPORTS) == set(names), set(L.EXPORTS) ^ set(names)


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.md2_version()
    assert isinstance(lib.md2_last_error(), bytes)


def test_desc_layout_matches_c():
    code = r'''
#include <stdio.h>
#include <stddef.h>
#include "md2.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(md2_vsl_desc), offsetof(md2_vsl_desc, disparity),
         offsetof(md2_vsl_desc, K), offsetof(md2_vsl_desc, automask), offsetof(md2_vsl_desc, loss),
         offsetof(md2_vsl_desc, grad_source), offsetof(md2_vsl_desc, saved), offsetof(md2_vsl_desc, zero_grad_source));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        vals = [int(v) for v in subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()]
    D = L.VslDesc
    assert vals == [C.sizeof(D), D.disparity.offset, D.K.offset, D.automask.offset, D.loss.offset,
                    D.grad_source.offset, D.saved.offset, D.zero_grad_source.offset]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    with pytest.raises(M.Md2Error):
        M.Context.get(torch.device("cuda", 0))
    x = torch.rand(1, 1, 8, 8)
    with pytest.raises(M.Md2Error):
        M.SSIM()(x, x)
    h = C.c_void_p()
    assert lib.md2_create(0, C.byref(h)) != 0
    assert b"no CUDA device" in lib.md2_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "monodepth2.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("# oracle", ""), f


def test_package_synthetic_data_equals_test_side_generator():
    """bench.py's CUDA arm builds its inputs with the package's generator, the CPU arm and the
    parity tests with the oracle-side one: both must produce identical data."""
    from monodepth2_jl_b200 import synthetic as SY
    from oracle import torch_oracle as O
    for kw in (dict(N=2, C=1, H=24, W=40, seed=42), dict(N=1, C=3, H=16, W=32, seed=7, full_res_disp=True, pose_sigma=0.1)):
        a, b = SY.synthetic_batch(**kw), O.synthetic_batch(**kw)
        assert torch.equal(a[0], b[0])
        for k in (1, 2, 3):
            assert all(torch.equal(u, v) for u, v in zip(a[k], b[k]))
    for u, v in zip(SY.make_K(416, 128), O.make_K(416, 128)):
        assert torch.equal(u, v)


def test_julia_binding_is_in_step_with_the_header():
    """static checks of julia/Monodepth2B200.jl (no Julia in the image): its mirror of md2_vsl_desc lists the header's
    fields in the header's order, every ccall names an exported symbol, and it extends the reference's generics instead
    of defining its own (`import ..Monodepth: ...`)"""
    import re
    hdr = open(os.path.join(ROOT, "include", "md2.h")).read()
    body = hdr[hdr.index("typedef struct md2_vsl_desc {"):hdr.index("} md2_vsl_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = []
    for stmt in body.split("{", 1)[1].split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        first, *rest = stmt.split(",")
        names = [first.split()[-1]] + [r.strip() for r in rest]
        c_fields += [re.sub(r"\[.*\]", "", n).lstrip("*") for n in names]
    jl = open(os.path.join(ROOT, "monodepth2.jl_b200", "julia", "Monodepth2B200.jl")).read()
    jbody = jl[jl.index("struct VslDesc"):]
    jbody = jbody[:jbody.index("\nend")]
    j_fields = re.findall(r"(\w+)::", re.sub(r"#.*", "", jbody))
    assert j_fields == c_fields
    for sym in set(re.findall(r"ccall\(\(:(\w+), LIB\)", jl)):
        assert sym in L.EXPORTS, sym
    assert "import ..Monodepth:" in jl and "module B200" in jl
    for name in ("train_loss", "slow_depth", "disparity_to_depth", "composeT", "smooth_loss", "prediction_loss", "automasking_loss"):
        assert re.search(r"import \.\.Monodepth:[^#]*\b" + name + r"\b", jl, flags=re.S), name
    assert not re.search(r"^struct (SSIM|Backproject|Project)\b", jl, flags=re.M)      # the reference's own structs are extended
