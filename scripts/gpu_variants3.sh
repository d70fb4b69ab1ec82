#!/bin/bash
# build-variant sweep over the C=3 configurations
for V in "$@"; do
  MD2_NVCC_EXTRA="$V" python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build(force=True)" 2> gpurun_out/variant_build.err || { echo "build failed: $V"; continue; }
  echo "[$V]"; python scripts/sweep_configs.py --c3 2>/dev/null | python -c '
import sys, json
for l in sys.stdin:
    d = json.loads(l); print("   ", d["W"], d["N"], d["automask"], d["ms_per_step"], d["march_kernel_ms"], d["kernel_frac_of_hbm_peak"])'
done
