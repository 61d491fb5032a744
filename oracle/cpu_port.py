"""ctypes front end of oracle/saa_oracle.c (TEST INFRA / CPU BASELINE ONLY).

``drone_assemble`` produces the same per-iteration values as the CUDA path
(u-column entries at their CSC positions, sample-row upper bounds, mean-row
sums); ``time_drone`` times it on the host cores for bench.py's ``cpu_baseline``.
"""
import ctypes as C
import os
import subprocess
import time

import numpy as np

from riskaversetrajopt_b200.drone import drone_params as dp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsaa_oracle.so")


class _Params(C.Structure):
    _fields_ = [("dt", C.c_double), ("beta", C.c_double), ("drag", C.c_double), ("kp", C.c_double),
                ("kv", C.c_double), ("x_init", C.c_double * 6), ("x_final", C.c_double * 6),
                ("obs_pos", (C.c_double * 2) * 3), ("mult", C.c_double), ("pad", C.c_double),
                ("escale", C.c_double)]


def _lib():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "saa_oracle.c")):
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    lib = C.CDLL(_SO)
    lib.saa_oracle_threads.restype = C.c_int
    lib.saa_oracle_drone_assemble.restype = None
    lib.saa_oracle_drone_assemble.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 10
    return lib


def drone_col_offsets(M, S=20):
    """Positions of sample 0's sub-run per (axis, j) in a buffer that holds ONLY the
    u-column block laid out as in the CSC matrix for M samples (final rows and
    control rows included, so offsets equal the real matrix's)."""
    off = np.zeros(2 * (S - 1), dtype=np.int64)
    pos = 0
    for j in range(S):
        for a in range(3):
            nfin = (1 if j <= S - 2 else 0) + 1
            run = 3 * (S - 1 - j) if (a < 2 and j <= S - 2) else 0
            if run:
                off[a * (S - 1) + j] = pos + nfin
            pos += nfin + M * run + 1
    return off, pos


def drone_assemble(us, masses, DWs, obs_Qs, S=20, mult=0.01, pad=0.0, escale=1.0, out=None):
    """``out``: (Ax, ub, sums, Z, off) of a previous call with the same M: the output buffers are
    reused (every entry the port owns is overwritten; the final-row / control-row slots stay 0)."""
    lib = _lib()
    M = masses.shape[0]
    if out is None:
        off, n = drone_col_offsets(M, S)
        Ax = np.zeros(n)
        ub = np.empty(M * 3 * S)
        sums = np.empty(3 * (S - 1) + 3 * S + 6)
        Z = np.empty(M)
    else:
        Ax, ub, sums, Z, off = out
    p = _Params()
    p.dt, p.beta, p.drag = dp.T / S, dp.beta, dp.drag_coefficient
    p.kp, p.kv = 0.05, 0.25
    p.x_init[:] = list(dp.x_init); p.x_final[:] = list(dp.x_final)
    for o in range(3):
        for a in range(2):
            p.obs_pos[o][a] = float(dp.obs_positions[o][a])
    p.mult, p.pad, p.escale = mult, pad, escale
    us = np.ascontiguousarray(us, dtype=np.float64)
    masses, DWs, obs_Qs = (np.ascontiguousarray(x, dtype=np.float64) for x in (masses, DWs, obs_Qs))
    lib.saa_oracle_drone_assemble(M, S, masses.ctypes.data, DWs.ctypes.data, obs_Qs.ctypes.data,
                                  us.ctypes.data, C.addressof(p), Ax.ctypes.data, off.ctypes.data,
                                  ub.ctypes.data, sums.ctypes.data, Z.ctypes.data)
    return Ax, ub, sums, Z, off


def synthetic_samples(M_s, S=20, seed=0):
    rs = np.random.RandomState(seed)
    masses = rs.uniform(dp.mass_nom - dp.mass_delta, dp.mass_nom + dp.mass_delta, M_s)
    obs_Qs = np.zeros((M_s, 3, 3, 3))
    for o in range(3):
        for d in range(3):
            obs_Qs[:, o, d, d] = 1. / (dp.obs_radii[o] + rs.uniform(-dp.obs_radii_deltas, dp.obs_radii_deltas, M_s))**2
    DWs = np.sqrt(dp.dt) * rs.randn(M_s, S, 6)
    return masses, DWs, obs_Qs


def time_drone(us, budget_s=20.0, M_s=None, S=20, max_reps=2000):
    """-> cpu_baseline dict for bench.py (all host threads OpenMP gives us).  The output buffers are
    allocated (and page-faulted) once, outside the timed calls, like the GPU arm's."""
    lib = _lib()
    cores = int(lib.saa_oracle_threads())
    M_s = int(M_s or 100_000)
    masses, DWs, obs_Qs = synthetic_samples(M_s, S)
    out = drone_assemble(us, masses, DWs, obs_Qs)        # warm-up (page faults, threads)
    drone_assemble(us, masses, DWs, obs_Qs, out=out)
    reps, t_total = 0, 0.0
    while t_total < budget_s and reps < max_reps:
        t0 = time.perf_counter()
        drone_assemble(us, masses, DWs, obs_Qs, out=out)
        t_total += time.perf_counter() - t0
        reps += 1
    dt_step = t_total / reps
    return {"value": M_s * S / dt_step, "unit": "samples*steps/s", "cores": cores, "kind": "port",
            "ms_per_step": dt_step * 1e3,
            "sample": f"oracle/saa_oracle.c (closed-form restatement writing the CSC values directly into host "
                      f"memory, C + OpenMP, {cores} threads) on {M_s} of the 10^6 samples, {reps} repetitions, "
                      f"output buffers reused"}
