#!/bin/bash
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_forced.py tests/test_gpu_edge.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 1000 --warmup 20 --no-train-step --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frames/s', d['value'], 'e2e', e['value'], 'sync', e['value_synchronous'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 30 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-train-step > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv
python - <<'PY'
import torch, time
dev = torch.device("cuda", 0)
for mb in (7.4, 2.3, 32.0):
    n = int(mb * 1e6 / 4)
    h = torch.zeros(n).pin_memory(); d = torch.empty(n, device=dev)
    for direction in ("h2d", "d2h"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
        print(f"{direction} {mb} MB pinned: {dt*1e6:.1f} us = {mb*1e-3/dt:.1f} GB/s")
PY
