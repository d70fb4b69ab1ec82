#!/bin/bash
# executed global reductions of the marching kernel at config 4, merged vs plain scatter
mkdir -p gpurun_out
for V in "-DMD2_M2_MERGE=0" "-DMD2_M2_MERGE=1"; do
  MD2_NVCC_EXTRA="$V" python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','monodepth2.jl_b200/build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build()" 2> gpurun_out/variant_build.err || { echo "build failed: $V"; continue; }
  echo "[$V]"
  MD2_NO_REPLAY=1 timeout 300 ncu --metrics l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sectors_op_red.sum --clock-control none -k regex:march2 -s 10 -c 1 --csv --log-file gpurun_out/red_count_${V: -1}.csv python bench.py --config 4 --steps 12 --warmup 3 --no-cpu-baseline --no-train-step > /dev/null 2>&1
done
