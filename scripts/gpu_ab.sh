#!/bin/bash
# A/B of older trees (exported under build/old_<commit>) against the current one on the C=3 sweep rows
for T in build/old_e1e98ab build/old_9b7c9d3 .; do
  echo "== $T"
  ( cd $T && python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build()" 2>/dev/null && python scripts/sweep_configs.py --c3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['W'], d['N'], d['C'], d['automask'], d['ms_per_step'], d['march_kernel_ms'], d['kernel_frac_of_hbm_peak'])" )
done
