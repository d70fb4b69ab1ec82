"""Import shim: the product package lives in the directory `monodepth2.jl_b200/` (a name
Python's import statement cannot spell), so `import monodepth2_jl_b200` redirects there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "monodepth2.jl_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
