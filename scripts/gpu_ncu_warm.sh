#!/bin/bash
# full ncu capture of the marching kernel with caches left alone between replay passes (warm L2)
TAG=${1:-warm}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:march -s 30 -c 1 -o gpurun_out/${TAG}_fused_warm \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_warm.log 2>&1; echo "ncu warm rc=$?"
