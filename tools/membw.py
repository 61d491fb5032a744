"""HBM bandwidth probes on this box: pure write (fill), pure read (sum), copy."""
import torch, numpy as np
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))
N = 1_200_000_000
x = torch.empty(N, dtype=torch.float64, device='cuda')
y = torch.empty(N, dtype=torch.float64, device='cuda')
for name, fn, nbytes in (("fill_ (write 9.6 GB)", lambda: x.fill_(1.5), 8*N), ("zero_ (memset 9.6 GB)", lambda: x.zero_(), 8*N),
                         ("copy_ (read+write 19.2 GB)", lambda: y.copy_(x), 16*N), ("sum (read 9.6 GB)", lambda: x.sum(), 8*N),
                         ("mul_ in place (r+w 19.2 GB)", lambda: x.mul_(1.0001), 16*N)):
    med, mn = t(fn)
    print(f"{name:32s} median {med:7.3f} ms  {nbytes/med/1e6:7.0f} GB/s   best {nbytes/mn/1e6:7.0f} GB/s")
