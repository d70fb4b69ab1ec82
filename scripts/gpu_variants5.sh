#!/bin/bash
# build-variant sweep of the marching kernel over configs 2, 3, 4
mkdir -p gpurun_out
for V in "$@"; do
  MD2_NVCC_EXTRA="$V" python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build()" 2> gpurun_out/variant_build.err || { echo "build failed: $V"; tail -5 gpurun_out/variant_build.err; continue; }
  echo "[$V]"
  for CFG in 2 3 4; do
    timeout 300 python bench.py --config $CFG --steps 300 --warmup 10 --no-train-step --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   config $CFG: ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])"
  done
  if [ "$V" != "-DMD2_M2_MERGE=0" ]; then timeout 600 python -m pytest tests/test_gpu_forced.py -m gpu -x -q 2>&1 | tail -1; fi
done
