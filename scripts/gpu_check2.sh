#!/bin/bash
# One GPU-box pass of round 2: all GPU tests, smoke, bench (default), launch list, full ncu capture of the marching kernel
TAG=${1:-r2y}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; echo "reference rc=$?"
MD2_NO_REPLAY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-train-step > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
MD2_NO_REPLAY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:march2 -s 30 -c 1 -o gpurun_out/${TAG}_fused \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-train-step > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
