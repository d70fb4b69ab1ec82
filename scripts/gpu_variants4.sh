#!/bin/bash
# build-variant sweep on the default bench (config 2): prints step / kernel times and the launch list
TAG=$1; shift
mkdir -p gpurun_out
for V in "$@"; do
  MD2_NVCC_EXTRA="$V" python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build()" 2> gpurun_out/variant_build.err || { echo "build failed: $V"; tail -5 gpurun_out/variant_build.err; continue; }
  echo "[$V]"
  timeout 300 python bench.py --steps 1000 --warmup 20 --no-train-step --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frames/s', d['value'], 'e2e', d['e2e']['value'], 'sync', d['e2e']['value_synchronous'])"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 30 --csv --log-file gpurun_out/${TAG}_launches_tmp.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-train-step > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/${TAG}_launches_tmp.csv | sed 's/^/   /'
done
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_forced.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3
