#!/bin/bash
# ncu --set full of the marching kernel of one bench step.  Usage: bash scripts/gpu_march_profile.sh <tag>
TAG=${1:-m}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 30 -c 1 -o gpurun_out/${TAG}_fused \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
