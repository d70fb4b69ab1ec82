#!/bin/bash
mkdir -p gpurun_out
MD2_NVCC_EXTRA="-DMD2_PDL_EARLY=0" python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build()" 2> gpurun_out/variant_build.err || { echo "build failed"; tail -5 gpurun_out/variant_build.err; exit 1; }
for P in 0 1 2 3; do MD2_PDL=$P python bench.py --steps 1000 --no-train-step --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('late trigger, MD2_PDL=$P ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], d['variants']['fwd_only_cold_ms'], d['variants']['fwdbwd_smooth_disparity_ms'])"; done
