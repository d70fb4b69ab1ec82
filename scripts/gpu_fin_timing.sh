#!/bin/bash
# where the finish kernel's time goes: builds that skip one part of it (-DMD2_FIN_TIMING=k, wrong results), kernel times under ncu
mkdir -p gpurun_out
B="import importlib.util; spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build()"
for V in "$@"; do
  MD2_NVCC_EXTRA="$V" python -c "$B" 2> gpurun_out/variant_build.err || { echo "build failed: $V"; tail -5 gpurun_out/variant_build.err; continue; }
  echo "[variant '$V']"
  MD2_NO_REPLAY=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/fin_ncu.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-train-step > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/fin_ncu.csv | tail -3
  timeout 200 python bench.py --steps 1000 --no-cpu-baseline --no-train-step 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   ms/step', d['ms_per_step'], d['roofline']['step_phases'])"
done
