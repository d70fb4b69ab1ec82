"""experiment: pinned host <-> device copy bandwidth and small-copy latency on the GPU box"""
import time
import torch
dev = torch.device("cuda", 0)
for mb in (1, 2, 8, 64):
    h = torch.empty(mb * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
    d = torch.empty_like(h, device=dev)
    for direction in ("h2d", "d2h"):
        for _ in range(3):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{direction} {mb:3d} MB: {ms * 1e3:8.1f} us  {mb / 1024 / (ms * 1e-3):6.1f} GB/s")
