"""Kernel micro-benchmark for build variants: python tools/kbench.py lib1.so lib2.so ...
Each library is timed in its own subprocess (drone assemble, M = 10^6, CUDA events)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, %r)
import bench
from riskaversetrajopt_b200 import _lib
from riskaversetrajopt_b200.device_path import DevicePath
from riskaversetrajopt_b200.drone import drone_params as dp
M = int(os.environ.get("KB_M", "1000000"))
dev = torch.device("cuda", 0)
if os.environ.get("KB_PROBLEM", "drone") == "car":
    from riskaversetrajopt_b200.car import driving_params as cp
    from riskaversetrajopt_b200.car.driving import BETA
    g = torch.Generator(device=dev); g.manual_seed(0)
    x0 = torch.as_tensor(cp.state_init, device=dev).repeat(M, 1)
    x0[:, 4:] += torch.randn((M, 4), generator=g, device=dev, dtype=torch.float64) * torch.tensor([0.1, 0.1, 1e-4, 1e-4], device=dev, dtype=torch.float64)
    ws = 0.025 + 0.15 * torch.rand(M, generator=g, device=dev, dtype=torch.float64)
    wr = 0.005 + 0.09 * torch.rand(M, generator=g, device=dev, dtype=torch.float64)
    DW = float(np.sqrt(cp.dt)) * torch.randn((M, 20, 8), generator=g, device=dev, dtype=torch.float64)
    path = DevicePath(_lib.SAA_CAR, 'saa', 20, 0.05, M, device=0)
    path.set_params_car(cp, BETA, cp.OSQP_TOL); path.set_samples_car(x0, ws, wr, DW)
    torch.cuda.synchronize(); del x0, ws, wr, DW; path._keep = []
    us = np.full((20, 2), 0.01) + 0.1 * np.random.RandomState(0).randn(20, 2)
else:
    DWs, masses, obs_Qs = bench.synthetic_drone_samples(M, 0, dev)
    path = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, M, device=0)
    path.set_params_drone(dp, dp.OSQP_TOL); path.set_samples_drone(masses, DWs, obs_Qs)
    torch.cuda.synchronize(); del DWs, masses, obs_Qs; path._keep = []
    us = bench.bench_us()
for _ in range(3): path.assemble(us, 2)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); path.assemble(us, 2, finalize=False); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
# checksum so that variants can be compared for equality
buf = path.buffers()
nv = (380 if os.environ.get("KB_PROBLEM", "drone") == "car" else 1140) * M
print("RESULT", float(np.median(ts)), float(min(ts)), float(buf['Ax'][:nv].sum().item()), float(buf['u'][:-200].sum().item()))
''' % ROOT

for lib in sys.argv[1:]:
    env = dict(os.environ, SAA_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    if line:
        _, med, mn, cs1, cs2 = line[0].split()
        bps = 3576 if os.environ.get("KB_PROBLEM", "drone") == "car" else 10144
        gbs = bps * int(os.environ.get("KB_M", "1000000")) / (float(med) * 1e-3) / 1e9
        print(f"{os.path.basename(lib):28s} median {float(med):7.3f} ms  min {float(mn):7.3f} ms  {gbs:7.1f} GB/s  "
              f"frac {gbs / 6555.8:.3f}  checksum {cs1} {cs2}", flush=True)
    else:
        print(os.path.basename(lib), "FAILED", out.stderr[-800:], flush=True)
