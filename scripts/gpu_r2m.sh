#!/bin/bash
# round-2 GPU pass: all GPU tests + bench with the training-step section
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 500 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
