"""Device-time every kernel family once (drone fp64/fp32, car, hopper, rollout/cvar): python tools/kbench_all.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from riskaversetrajopt_b200 import _lib
from riskaversetrajopt_b200.device_path import DevicePath
from riskaversetrajopt_b200.drone import drone_params as dp
from riskaversetrajopt_b200.car import driving_params as cp
from riskaversetrajopt_b200.car.driving import BETA

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

dev = torch.device("cuda", 0)
M = int(os.environ.get("KB_M", "1000000"))
PEAK = 6555.8
ONLY = os.environ.get("KB_ONLY", "drone,car,hopper").split(",")
for prec, bps in (("fp64", 10144), ("fp32", 5072)) if "drone" in ONLY else ():
    DWs, masses, obs_Qs = bench.synthetic_drone_samples(M, 0, dev)
    p = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, M, device=0, precision=prec)
    p.set_params_drone(dp, dp.OSQP_TOL); p.set_samples_drone(masses, DWs, obs_Qs)
    torch.cuda.synchronize(); del DWs, masses, obs_Qs; p._keep = []
    us = bench.bench_us()
    t = timeit(lambda: p.assemble(us, 2, finalize=False))
    print(f"drone assemble {prec}: {t:.3f} ms  {M*bps/t/1e6:.0f} GB/s  frac {M*bps/t/1e6/PEAK:.3f}")
    if prec == "fp64":
        t = timeit(lambda: p.cvar_terms(us, 0.0, 1e-6, want_Z=True))
        print(f"drone cvar_terms: {t:.3f} ms  ({M*(536+8)/t/1e6:.0f} GB/s algorithmic)")
        t = timeit(lambda: p.rollout(us))
        print(f"drone rollout   : {t:.3f} ms  ({M*(536+126*8)/t/1e6:.0f} GB/s algorithmic)")
    del p; torch.cuda.empty_cache()
# car
if "car" in ONLY:
    g = torch.Generator(device=dev); g.manual_seed(0)
    x0 = torch.as_tensor(cp.state_init, device=dev).repeat(M, 1)
    x0[:, 4:] += torch.randn((M, 4), generator=g, device=dev, dtype=torch.float64) * torch.tensor([0.1, 0.1, 1e-4, 1e-4], device=dev, dtype=torch.float64)
    ws = 0.025 + 0.15 * torch.rand(M, generator=g, device=dev, dtype=torch.float64)
    wr = 0.005 + 0.09 * torch.rand(M, generator=g, device=dev, dtype=torch.float64)
    DW = float(np.sqrt(cp.dt)) * torch.randn((M, 20, 8), generator=g, device=dev, dtype=torch.float64)
    for prec, bps in (("fp64", 3576), ("fp32", 1788)):
        p = DevicePath(_lib.SAA_CAR, 'saa', 20, 0.05, M, device=0, precision=prec)
        p.set_params_car(cp, BETA, cp.OSQP_TOL); p.set_samples_car(x0, ws, wr, DW)
        torch.cuda.synchronize(); p._keep = []
        usc = np.full((20, 2), 0.01) + 0.1 * np.random.RandomState(0).randn(20, 2)
        t = timeit(lambda: p.assemble(usc, 2, finalize=False))
        print(f"car assemble {prec}: {t:.3f} ms  {M*bps/t/1e6:.0f} GB/s  frac {M*bps/t/1e6/PEAK:.3f}  ({M*20/t/1e6:.2f} G sample-steps/s)")
        del p; torch.cuda.empty_cache()
    del x0, ws, wr, DW; torch.cuda.empty_cache()
# hopper
if "hopper" in ONLY:
    from riskaversetrajopt_b200.hopper import hopper as hp
    Mh = min(M, 1000000)
    rs = np.random.RandomState(0)
    f = (0.025 * np.sqrt(2 / 30) * rs.uniform(0, 1, (Mh, 30)), rs.uniform(0, np.pi, (Mh, 30)), rs.uniform(0, 2 * np.pi, (Mh, 30)))
    px = np.linspace(0, 0.2, 20)
    for prec in ("fp64", "fp32"):
        m = hp.Model(Mh, 'saa', 0.1, f, precision=prec)
        lam = torch.randn(Mh * 20, device=dev, dtype=torch.float64)
        hs = torch.zeros(40, device=dev, dtype=torch.float64)
        def hop(hess=False):
            _lib.check(_lib.lib.saa_hopper_friction(m._h, 20, px.ctypes.data, m._mu.data_ptr(), m._dmu.data_ptr(),
                                                    lam.data_ptr() if hess else None, hs.data_ptr() if hess else None, m._stream()), m._h)
        t = timeit(hop)
        print(f"hopper friction {prec} (g, jac): {t:.3f} ms for M={Mh}: {Mh*600/t/1e6:.1f} G sincos/s, {Mh*1520/t/1e6:.0f} GB/s algorithmic")
        t = timeit(lambda: hop(True))
        print(f"hopper friction {prec} (+ Hessian sums): {t:.3f} ms for M={Mh}: {Mh*600/t/1e6:.1f} G sincos/s")
        Z = rs.uniform(-1, 1, hp.num_vars(Mh)) if prec == "fp64" else Z
        import ctypes as C
        pt = m._point(Z)
        y = torch.as_tensor(Z[(hp.S + 1) * 8 + hp.S * 4:-2].copy()).to(dev)
        gbuf = torch.empty(m.n_rows, dtype=m._jac_dev.dtype, device=dev)
        def g_and_jac():
            _lib.check(_lib.lib.saa_hopper_g(m._h, C.byref(pt), y.data_ptr(), gbuf.data_ptr(), m._stream()), m._h)
            _lib.check(_lib.lib.saa_hopper_jac(m._h, C.byref(pt), m._jac_dev.data_ptr(), m._stream()), m._h)
        t = timeit(g_and_jac)
        es = 8 if prec == "fp64" else 4
        print(f"hopper saa_hopper_g + saa_hopper_jac {prec}: {t:.3f} ms for M={Mh} (two launches, {2*Mh*600/t/1e6:.1f} G sincos/s; "
              f"{Mh*(2*720+100*es)/t/1e6:.0f} GB/s: 720 B read per launch, {100*es} B written per sample)")
        t = timeit(lambda: _lib.check(_lib.lib.saa_hopper_jac(m._h, C.byref(pt), m._jac_dev.data_ptr(), m._stream()), m._h))
        print(f"hopper saa_hopper_jac {prec}: {t:.3f} ms ({Mh*600/t/1e6:.1f} G sincos/s)")
        del m
