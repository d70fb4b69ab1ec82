#!/bin/bash
# e2e: image groups per call in the double-buffered host path
mkdir -p gpurun_out
for G in 1 2 4; do
  timeout 300 python bench.py --steps 500 --warmup 20 --no-train-step --no-cpu-baseline --e2e-groups $G 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('groups $G: pipelined', e['value'], 'ms', e['ms_per_step'], 'sync', e['value_synchronous'], 'g1', e['value_g1'], '| step ms', d['ms_per_step'])"
done
