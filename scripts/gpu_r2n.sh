#!/bin/bash
# 2-GPU pass: NCCL test of the data-parallel step + bench at N=2 (training-step section with the all-reduce)
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_optim.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 20 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_2gpu.err
cat gpurun_out/${TAG}_bench_2gpu.json
