#!/bin/bash
# automask pre-pass (photometric_min forward): thread-staged tiles (-DMD2_PM_BULK=0) against copy-engine staging (default)
mkdir -p gpurun_out
B="import importlib.util; spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build()"
for V in "-DMD2_PM_BULK=0" ""; do
  MD2_NVCC_EXTRA="$V" python -c "$B" 2> gpurun_out/variant_build.err || { echo "build failed: $V"; tail -5 gpurun_out/variant_build.err; continue; }
  echo "[variant '$V']"
  python scripts/exp/pm_bench.py 2>&1 | tail -4
  timeout 300 python bench.py --config 3 --steps 300 --no-cpu-baseline --no-train-step 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   config 3: ms/step', d['ms_per_step'], 'frames/s', d['value'])"
done
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fuzz.py tests/test_gpu_forced.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -3
