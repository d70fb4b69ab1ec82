#!/bin/bash
# quick round-2 GPU pass: fused parity tests + bench (+ optional full ncu capture of the marching kernel)
# usage: bash scripts/gpu_r2q.sh <tag> [ncu]
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fused.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("ms/step", d["ms_per_step"], "frames/s", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
PY
if [ "$2" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march -s 30 -c 1 -o gpurun_out/${TAG}_fused \
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
