"""Device-to-host bandwidth into pinned memory: one stream vs several (chunks in flight on
different streams), for the e2e path of bench.py.  python tools/d2h_bw.py"""
import time
import torch
n = 1_200_000_000            # doubles = 9.6 GB
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
def run(nstreams, chunks):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step = n // chunks
    for c in range(chunks):
        with torch.cuda.stream(streams[c % nstreams]):
            h[c * step:(c + 1) * step].copy_(d[c * step:(c + 1) * step], non_blocking=True)
    torch.cuda.synchronize()
    return n * 8 / (time.perf_counter() - t0) / 1e9
for ns, ch in ((1, 1), (1, 8), (2, 2), (2, 8), (4, 4), (4, 16)):
    run(ns, ch)
    print(f"streams {ns}  chunks {ch:2d}: {max(run(ns, ch) for _ in range(3)):6.1f} GB/s")
