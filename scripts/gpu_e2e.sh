#!/bin/bash
python scripts/exp/pcie_probe.py
for G in 1 2 4 8; do
  echo "groups=$G $(python bench.py --steps 300 --warmup 5 --no-cpu-baseline --e2e-groups $G 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["autograd_api_value"])')"
done
