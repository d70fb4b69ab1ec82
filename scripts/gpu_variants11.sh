#!/bin/bash
mkdir -p gpurun_out
for V in "$@"; do
  MD2_NVCC_EXTRA="$V" python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build()" 2> gpurun_out/variant_build.err || { echo "build failed: $V"; tail -5 gpurun_out/variant_build.err; continue; }
  echo "[$V]"
  for i in 1 2; do timeout 300 python bench.py --steps 1000 --no-train-step --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); v=d['variants']; print('   ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'fwd-only', v['fwd_only_cold_ms'], 'smooth', v['march_kernel_smooth_disparity_ms'])"; done
  timeout 600 python -m pytest tests/test_gpu_forced.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -1
done
