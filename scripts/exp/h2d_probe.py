import torch, time
dev = torch.device("cuda", 0)
torch.cuda.init(); torch.zeros(1, device=dev)
def rate(mb, nstreams=1, direction="h2d", reps=60):
    n = int(mb * 1e6 / 4) // nstreams
    hs = [torch.empty(n).pin_memory() for _ in range(nstreams)]
    ds = [torch.empty(n, device=dev) for _ in range(nstreams)]
    ss = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    def burst(k):
        for _ in range(k):
            for h, d, s in zip(hs, ds, ss):
                with torch.cuda.stream(s):
                    (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        for s in ss: s.synchronize()
    burst(10)
    t0 = time.perf_counter(); burst(reps); dt = (time.perf_counter() - t0) / reps
    return dt * 1e6, mb * 1e-3 / dt
for mb in (2.3, 4.0, 5.1, 6.0, 7.4, 8.0, 10.0, 16.0, 32.0):
    a = rate(mb, 1); b = rate(mb, 2); c = rate(mb, 4)
    print(f"h2d {mb:5.1f} MB: 1 stream {a[0]:7.1f} us {a[1]:5.1f} GB/s | 2 streams {b[0]:7.1f} us {b[1]:5.1f} GB/s | 4 streams {c[0]:7.1f} us {c[1]:5.1f} GB/s")
for mb in (2.3, 7.4):
    a = rate(mb, 1, "d2h"); print(f"d2h {mb:5.1f} MB: 1 stream {a[0]:7.1f} us {a[1]:5.1f} GB/s")
