#!/bin/bash
# build-variant sweep of the marching kernel: each argument is a quoted MD2_NVCC_EXTRA string
mkdir -p gpurun_out
for V in "$@"; do
  MD2_NVCC_EXTRA="$V" python -c "
import importlib.util,sys
spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build(force=True)" 2> gpurun_out/variant_build.err || { echo "build failed: $V"; tail -3 gpurun_out/variant_build.err; continue; }
  echo "[$V] $(python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])')"
done
