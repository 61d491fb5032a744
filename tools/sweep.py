"""BASELINE config 4 on one GPU: drone SAA linearize+assemble for M = 1e4 ... 1e7 (FP64 and FP32),
device-timed.  python tools/sweep.py [Mmax]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from riskaversetrajopt_b200 import _lib
from riskaversetrajopt_b200.device_path import DevicePath
from riskaversetrajopt_b200.drone import drone_params as dp

dev = torch.device("cuda", 0)
Mmax = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
PEAK = 6555.8
us = bench.bench_us()
print(f"{'M':>10s} {'prec':>5s} {'ms':>9s} {'samples*steps/s':>17s} {'GB/s':>8s} {'of peak':>8s} {'values GB':>10s}")
for M in (10_000, 100_000, 1_000_000, 10_000_000):
    if M > Mmax:
        break
    for prec, bps in (("fp64", 10144), ("fp32", 5072)):
        need = (1263 * M + 2 * 61 * M) * (8 if prec == "fp64" else 4) + 536 * M + 1200 * M
        if need > 0.9 * torch.cuda.get_device_properties(0).total_memory:
            print(f"{M:10d} {prec:>5s}   skipped (needs {need/1e9:.0f} GB)")
            continue
        DWs, masses, obs_Qs = bench.synthetic_drone_samples(M, 0, dev)
        p = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, M, device=0, precision=prec)
        p.set_params_drone(dp, dp.OSQP_TOL); p.set_samples_drone(masses, DWs, obs_Qs)
        torch.cuda.synchronize(); del DWs, masses, obs_Qs; p._keep = []; torch.cuda.empty_cache()
        for _ in range(3):
            p.assemble(us, 2)
        torch.cuda.synchronize()
        n = 20 if M <= 1_000_000 else 5
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); p.assemble(us, 2); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = float(np.median(ts))
        gbs = M * bps / t / 1e6
        print(f"{M:10d} {prec:>5s} {t:9.3f} {M*20/(t*1e-3):17.4g} {gbs:8.0f} {gbs/PEAK:8.3f} {1140*M*(8 if prec=='fp64' else 4)/1e9:10.2f}", flush=True)
        p.close(); del p; torch.cuda.empty_cache()
