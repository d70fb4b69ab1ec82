#!/bin/bash
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 600 python scripts/microbench_ops.py > gpurun_out/${TAG}_ops_microbench.jsonl 2> gpurun_out/${TAG}_ops_microbench.err; echo "microbench rc=$?"; tail -3 gpurun_out/${TAG}_ops_microbench.err
python -c "
import json
for l in open('gpurun_out/${TAG}_ops_microbench.jsonl'):
    d=json.loads(l); print(f\"{d['op'][:44]:44s} {d['W']}x{d['H']}x{d['N']:<3d} warm {d['warm_ms']:8.4f} cold {d['cold_ms']:8.4f} ms  {d['cold_GBps']:7.1f} GB/s {d['cold_frac_of_hbm_peak']:.3f}\")"
timeout 600 python bench.py --config 1 --steps 10 > gpurun_out/${TAG}_bench_c1.json 2> gpurun_out/${TAG}_bench_c1.err; echo "bench c1 rc=$?"; tail -3 gpurun_out/${TAG}_bench_c1.err; cat gpurun_out/${TAG}_bench_c1.json | cut -c1-1500
timeout 600 python scripts/sweep_configs.py > gpurun_out/${TAG}_config_sweep.jsonl 2>/dev/null; echo "sweep rc=$?"
python -c "
import json
for l in open('gpurun_out/${TAG}_config_sweep.jsonl'):
    d=json.loads(l); print(d['W'], d['H'], d['N'], d['C'], d['automask'], d['ms_per_step'], d['march_kernel_ms'], d['kernel_frac_of_hbm_peak'])"
