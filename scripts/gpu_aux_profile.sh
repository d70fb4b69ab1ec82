#!/bin/bash
# ncu --set full of the auxiliary kernels (prep, finish) of one bench step
TAG=${1:-aux}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'prep|finish_kernel' -s 20 -c 2 -o gpurun_out/${TAG}_aux \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-train-step > gpurun_out/${TAG}_ncu_aux.log 2>&1; echo "ncu aux rc=$?"
