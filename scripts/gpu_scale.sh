#!/bin/bash
# scaling run: bench at N GPUs (torchrun), loss path + training-step section
N=$1; TAG=${2:-r2v}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 1000 --warmup 20 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_${N}gpu.json')); print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'sync', d['e2e']['value_synchronous']); print(json.dumps(d['train_step']))"
