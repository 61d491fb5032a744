"""Where the end-to-end step of bench.py spends its time: kernel, the three device-to-host copies
(CUDA events), and the copy bandwidth for differently filled / allocated buffers.
python tools/e2e_diag.py"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from riskaversetrajopt_b200 import _lib
from riskaversetrajopt_b200.device_path import DevicePath
from riskaversetrajopt_b200.drone import drone_params as dp
M = 1_000_000
dev = torch.device("cuda", 0)
DWs, masses, obs_Qs = bench.synthetic_drone_samples(M, 0, dev)
path = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, M, device=0)
path.set_params_drone(dp, dp.OSQP_TOL); path.set_samples_drone(masses, DWs, obs_Qs)
path.set_output_geometry(M, 0)
torch.cuda.synchronize(); del DWs, masses, obs_Qs; path._keep = []
us = bench.bench_us()
n_rows, n_cols, nnz = path.pattern_sizes()
n_var = 1140 * M + 177
b = path.buffers()
hAx = torch.empty(n_var, dtype=torch.float64, pin_memory=True)
hu = torch.empty(n_rows, dtype=torch.float64, pin_memory=True)
hl = torch.empty(6, dtype=torch.float64, pin_memory=True)
st = torch.cuda.current_stream()
def step(tag):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    t0 = time.perf_counter()
    ev[0].record(); path.assemble(us, 2, finalize=True)
    ev[1].record(); hAx.copy_(b['Ax'][:n_var], non_blocking=True)
    ev[2].record(); hu.copy_(b['u'], non_blocking=True)
    ev[3].record(); hl.copy_(b['l'][:6], non_blocking=True)
    ev[4].record(); st.synchronize()
    t1 = time.perf_counter()
    print(tag, "wall %.1f ms" % ((t1 - t0) * 1e3), ["%.2f" % ev[i].elapsed_time(ev[i + 1]) for i in range(4)],
          "Ax GB/s %.1f  u GB/s %.1f" % (n_var * 8 / ev[1].elapsed_time(ev[2]) / 1e6, n_rows * 8 / ev[2].elapsed_time(ev[3]) / 1e6))
for i in range(4): step(i)
# host buffer touched by the CPU between iterations (as a solver would) ?
hAx.numpy()[::4096] += 0.0
step("after host touch")

def bw(dst, src, tag):
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 0
    for _ in range(3):
        a.record(); dst.copy_(src, non_blocking=True); c.record(); torch.cuda.synchronize()
        best = max(best, src.numel() * 8 / a.elapsed_time(c) / 1e6)
    print(f"{tag:60s} {best:6.1f} GB/s")
Ax = b['Ax']
bw(hAx, Ax[:n_var], "assembled Ax -> hAx")
fresh = torch.empty(n_var, dtype=torch.float64, device=dev)
bw(hAx, fresh, "fresh device tensor (uninitialised) -> hAx")
fresh.copy_(Ax[:n_var]); torch.cuda.synchronize()
bw(hAx, fresh, "fresh device tensor holding a copy of Ax -> hAx")
fresh.fill_(1.5); torch.cuda.synchronize()
bw(hAx, fresh, "fresh device tensor filled with a constant -> hAx")
fresh.normal_(); torch.cuda.synchronize()
bw(hAx, fresh, "fresh device tensor filled with random doubles -> hAx")
del fresh
h2 = torch.empty(1_200_000_000, dtype=torch.float64, pin_memory=True)
bw(h2[:n_var], Ax[:n_var], "assembled Ax -> another pinned buffer (1.2e9 doubles)")
