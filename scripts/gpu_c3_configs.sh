#!/bin/bash
# C=3 configurations: bench lines + one full ncu capture of the marching kernel at config 4
TAG=${1:-r2s}
mkdir -p gpurun_out
for CFG in 3 4; do
  timeout 600 python bench.py --config $CFG --steps 300 --warmup 10 --no-train-step --no-cpu-baseline > gpurun_out/${TAG}_bench_c${CFG}.json 2> gpurun_out/${TAG}_bench_c${CFG}.err; echo "bench c$CFG rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_c${CFG}.json')); print('   ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'frames/s', d['value'], 'e2e', d['e2e']['value'], 'sync', d['e2e']['value_synchronous'])"
done
MD2_NO_REPLAY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:march2 -s 12 -c 1 -o gpurun_out/${TAG}_c4_fused \
  python bench.py --config 4 --steps 12 --warmup 3 --no-cpu-baseline --no-train-step > gpurun_out/${TAG}_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
MD2_NO_REPLAY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 12 --csv --log-file gpurun_out/${TAG}_launches_c4.csv python bench.py --config 4 --steps 12 --warmup 3 --no-cpu-baseline --no-train-step > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches_c4.csv
