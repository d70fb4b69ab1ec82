#!/bin/bash
mkdir -p gpurun_out
for E in 0 1 2 4 8 16 32 63; do
  MD2_EXP=$E ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/exp_$E.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  echo "MD2_EXP=$E"; python scripts/launch_summary.py gpurun_out/exp_$E.csv | grep -v march
done
