#!/bin/bash
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 1000 --warmup 20 --no-train-step --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("ms/step", d["ms_per_step"], "frames/s", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"])
print({k: v for k, v in d["e2e"].items() if k not in ("api", "note")})
PY
MD2_NO_REPLAY=1 timeout 300 python bench.py --steps 1000 --warmup 20 --no-train-step --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no-replay: ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"
MD2_BENCH_G0=1 timeout 300 python bench.py --steps 1000 --warmup 20 --no-train-step --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('g=0: ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 30 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-train-step > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
MD2_NO_REPLAY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:march2 -s 30 -c 1 -o gpurun_out/${TAG}_fused \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-train-step > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
