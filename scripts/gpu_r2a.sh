#!/bin/bash
# round 2 pass: parity tests, bench of the single-warp kernel vs the two-warp kernel, launch list, full capture
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
MD2_MARCH_V1=1 timeout 600 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/${TAG}_bench_v1.json 2> gpurun_out/${TAG}_bench_v1.err; echo "bench v1 rc=$?"
cat gpurun_out/${TAG}_bench_v1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march -s 30 -c 1 -o gpurun_out/${TAG}_fused \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
