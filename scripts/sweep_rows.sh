#!/bin/bash
# bench the marching kernel at several chunk heights (MD2_MARCH_ROWS) -- tuning aid
for R in ${@:-16 22 26 32 43 64}; do
  echo "R=$R $(MD2_MARCH_ROWS=$R python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])')"
done
