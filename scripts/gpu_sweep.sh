#!/bin/bash
# bench at several chunk heights of the marching kernel (MD2_MARCH_ROWS); MD2_DEBUG prints the resident blocks/SM
for R in ${@:-0 64 43 32 26}; do
  echo "R=$R $(MD2_DEBUG=1 MD2_MARCH_ROWS=$R python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>gpurun_out/sweep_$R.err | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])') $(grep -m2 -E "resident|chunk height" gpurun_out/sweep_$R.err | tr "\n" " ")"
done
