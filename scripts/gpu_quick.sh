#!/bin/bash
# quick GPU pass: parity tests, bench, ncu launch list (no full capture).  Usage: bash scripts/gpu_quick.sh <tag> [pytest-args]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("ms/step", d["ms_per_step"], "frames/s", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv
