#!/bin/bash
mkdir -p gpurun_out
for V in "$@"; do
  MD2_NVCC_EXTRA="$V" python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build()" 2> gpurun_out/variant_build.err || { echo "build failed: $V"; tail -5 gpurun_out/variant_build.err; continue; }
  echo "[$V]"
  for CFG in 2 4; do timeout 300 python bench.py --config $CFG --steps 300 --no-train-step --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   config $CFG ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'])"; done
done
