#!/bin/bash
for NG in 0 1; do for G in 1 2 4; do
  echo "nograph=$NG groups=$G $( (if [ $NG = 1 ]; then export MD2_HOST_NO_GRAPH=1; fi; python bench.py --steps 300 --warmup 5 --no-cpu-baseline --e2e-groups $G 2>/dev/null) | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["autograd_api_value"])')"
done; done
