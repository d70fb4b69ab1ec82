#!/bin/bash
# build-variant x chunk-height sweep: args = pairs "NVCC_EXTRA|R"
mkdir -p gpurun_out
for V in "$@"; do
  X="${V%%|*}"; R="${V##*|}"
  MD2_NVCC_EXTRA="$X" python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build(force=True)" 2> gpurun_out/variant_build.err || { echo "build failed: $X"; continue; }
  echo "[$X R=$R] $(MD2_DEBUG=1 MD2_MARCH_ROWS=$R python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>gpurun_out/v.err | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])') $(grep -m1 resident gpurun_out/v.err | cut -c1-70)"
done
