"""BASELINE.json's full size (drone, M = 10^6 on one GPU) through size-independent properties,
evaluated on the device (the matrix values alone are 9.6 GB):
  * sampled samples' sub-runs of every u column and their bounds against the oracle;
  * the tail-reduced matrix (K = 125 000, another tiling / another handle / gathered inputs) is
    bitwise the full matrix restricted to the selected samples, in all 38 sample-carrying columns;
  * the means-only pass reproduces the assemble launch's expectation sums, Z_i agrees between the
    assemble, means and CVaR kernels; two launches are bitwise identical."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S = 20


def _col(j, a, M):
    """start of sample 0's run of u column (j, a), run length per sample (csrc/drone_kernels.cuh,
    DroneChain; tests/test_pattern.py checks the same layout against SciPy's)."""
    L = S - 1 - j
    return (9 * j + 2 + 3 * a) + M * (6 * j * (S - 1) - 3 * j * (j - 1) + 3 * a * L), 3 * L


def test_drone_million_samples_properties():
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~25 GB of device memory")
    sys.path.insert(0, ROOT)
    import bench
    from oracle.oracle_b import DroneOracleB
    from riskaversetrajopt_b200 import _lib
    from riskaversetrajopt_b200._lib import lib, check
    from riskaversetrajopt_b200.device_path import DevicePath
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.tail import TailSubproblem
    M = 1_000_000
    dev = torch.device("cuda", 0)
    DWs, masses, obs_Qs = bench.synthetic_drone_samples(M, 11, dev)
    path = DevicePath(_lib.SAA_DRONE, 'saa', S, 0.1, M, device=0)
    path.set_params_drone(dp, dp.OSQP_TOL)
    path.set_samples_drone(masses, DWs, obs_Qs)
    us = bench.bench_us()
    Z = torch.empty(M, dtype=torch.float64, device=dev)
    b = path.assemble(us, 2, Z=Z)
    Ax, u = b['Ax'], b['u']
    assert Ax.numel() == 1263 * M + 180
    sums = path.mean_sums.clone()
    row_s0 = 6 + 1 + M

    # ---- sampled samples against the oracle ------------------------------------------------
    idx = np.unique(np.concatenate([[0, 1, 15, 16, 17, M - 1, M - 16, M - 17],
                                    np.random.RandomState(0).randint(0, M, 48)]))
    it = torch.as_tensor(idx, device=dev)
    ref = DroneOracleB(S, DWs[it].cpu().numpy(), masses[it].cpu().numpy(), obs_Qs[it].cpu().numpy(), 'saa', 0.1)
    _, _, _, g_du, g_up = ref.per_sample(us)
    rel = lambda a, w: float(np.max(np.abs(a - w) / np.maximum(np.abs(w), 1e-12)))
    ub = u[(row_s0 + it[:, None] * 60 + torch.arange(60, device=dev)[None, :])].cpu().numpy()
    assert rel(ub, 0.01 * g_up.reshape(len(idx), -1)) < 1e-9
    for j in range(S - 1):
        for a in (0, 1):
            start, LEN = _col(j, a, M)
            got = Ax[start + it[:, None] * LEN + torch.arange(LEN, device=dev)[None, :]].cpu().numpy()
            want = 0.01 * g_du[:, :, j + 1:, j * 3 + a].reshape(len(idx), -1)
            assert rel(got, want) < 1e-9, (j, a)

    # ---- tail-reduced matrix == full matrix restricted (bitwise), all columns -----------------
    tail = TailSubproblem(path, margin=0.25)
    K = tail.K
    assert K == 125_000
    br = tail.assemble(us, 2)
    sel = tail.idx
    Zs, order = torch.sort(Z, descending=True, stable=True)
    assert torch.equal(torch.sort(order[:K]).values, sel)            # the K largest, ties to the smaller index
    ar = torch.arange(K, device=dev)
    for j in range(S - 1):
        for a in (0, 1):
            sf, LEN = _col(j, a, M)
            sr, _ = _col(j, a, K)
            e = torch.arange(LEN, device=dev)[None, :]
            assert torch.equal(br['Ax'][sr + ar[:, None] * LEN + e], Ax[sf + sel[:, None] * LEN + e]), (j, a)
    e = torch.arange(60, device=dev)[None, :]
    assert torch.equal(br['u'][6 + 1 + K + ar[:, None] * 60 + e], u[row_s0 + sel[:, None] * 60 + e])
    assert torch.allclose(br['l'][:6], b['l'][:6], rtol=1e-12, atol=1e-15)
    assert tail.left_out_margin(float(Zs[K - 1])) <= 0.0

    # ---- kernels agree with each other; launches are deterministic ------------------------------
    assert torch.allclose(path.mean_sums, sums, rtol=1e-12, atol=1e-9)   # means-only pass (run by tail.assemble)
    assert torch.allclose(tail.Z, Z, rtol=0, atol=1e-15)
    Zc, out3 = path.cvar_terms(us, t_risk=-0.5, sat_tol=1e-6)
    assert torch.allclose(Z - dp.OSQP_TOL, Zc, rtol=0, atol=1e-15)
    assert abs(out3[0].item() - torch.clamp(Zc + 0.5, min=0).sum().item()) <= 1e-9 * abs(out3[0].item())
    assert out3[2].item() == Zc.max().item()
    keep = Ax[:1140 * M + 177].clone()
    path.assemble(us, 2, Z=Z)
    assert torch.equal(keep, Ax[:1140 * M + 177])


def _car_col(j, c, M):
    """start of sample 0's run of u column (j, c) of the car matrix, run length (CarCol)."""
    L = S - 1 - j
    return (8 * j + 3 + 4 * c) + M * (2 * j * (S - 1) - j * (j - 1) + c * L), L


def test_car_million_samples_properties():
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~12 GB of device memory")
    from oracle.oracle_b import CarOracleB
    from riskaversetrajopt_b200.car.driving import Model, sample_uncertain_parameters
    M = 1_000_000
    np.random.seed(4)
    samples = sample_uncertain_parameters(M, 'saa')        # host RNG in the reference's order (~1.4 GB)
    model = Model(M, 'saa', 0.05, samples=samples)
    path = model.path
    dev = path.device
    us = model.initial_guess_us_mat() + 0.1 * np.random.RandomState(0).randn(S, 2)
    Z = torch.empty(M, dtype=torch.float64, device=dev)
    b = path.assemble(us, 2, Z=Z)
    Ax, u = b['Ax'], b['u']
    row_s0 = 4 + 1 + M
    # sampled samples against the oracle (generic dense 8x8 recursion)
    idx = np.unique(np.concatenate([[0, 1, 15, 16, M - 1, M - 17], np.random.RandomState(1).randint(0, M, 42)]))
    it = torch.as_tensor(idx, device=dev)
    ref = CarOracleB(*(a[idx] for a in samples), 'saa', 0.05)
    _, _, _, g_du, g_up, _ = ref.per_sample(us)
    rel = lambda a, w: float(np.max(np.abs(a - w) / np.maximum(np.abs(w), 1e-12)))
    ub = u[row_s0 + it[:, None] * S + torch.arange(S, device=dev)[None, :]].cpu().numpy()
    assert rel(ub, g_up) < 1e-9
    for j in range(S - 1):
        for c in (0, 1):
            start, L = _car_col(j, c, M)
            got = Ax[start + it[:, None] * L + torch.arange(L, device=dev)[None, :]].cpu().numpy()
            assert rel(got, g_du[:, j + 1:, j * 2 + c]) < 1e-9, (j, c)
    # tail-reduced matrix == full matrix restricted, bitwise
    tail = model.tail_subproblem(margin=0.25)
    K = tail.K
    br = tail.assemble(us, 2)
    sel, ar = tail.idx, torch.arange(K, device=dev)
    order = torch.sort(Z, descending=True, stable=True).indices
    assert torch.equal(torch.sort(order[:K]).values, sel)
    for j in range(S - 1):
        for c in (0, 1):
            sf, L = _car_col(j, c, M)
            sr, _ = _car_col(j, c, K)
            e = torch.arange(L, device=dev)[None, :]
            assert torch.equal(br['Ax'][sr + ar[:, None] * L + e], Ax[sf + sel[:, None] * L + e]), (j, c)
    e = torch.arange(S, device=dev)[None, :]
    assert torch.equal(br['u'][4 + 1 + K + ar[:, None] * S + e], u[row_s0 + sel[:, None] * S + e])
    assert torch.equal(br['l'][:4], b['l'][:4]) and torch.equal(br['u'][:4], b['u'][:4])
    # kernels agree with each other
    from riskaversetrajopt_b200.car import driving_params as cp
    Zc, out3 = path.cvar_terms(us, t_risk=-1.0, sat_tol=1e-6)
    assert torch.allclose(Z - cp.OSQP_TOL, Zc, rtol=0, atol=1e-14)
    assert out3[2].item() == Zc.max().item()


def test_drone_two_million_samples_int64_offsets():
    """nnz = 1263 M + 180 exceeds 2^31 from M ~ 1.7e6: 64-bit CSC offsets in the kernels and an int64
    pattern on the host.  M = 2e6 (20 GB of matrix values): sampled samples -- the first, the last
    and those around the 2^31 element boundary of every long column -- against the oracle, at
    positions taken from the int64 column pointers of the library."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs ~45 GB of device memory")
    sys.path.insert(0, ROOT)
    import bench
    from oracle.oracle_b import DroneOracleB
    from riskaversetrajopt_b200 import _lib
    from riskaversetrajopt_b200._lib import lib, check
    from riskaversetrajopt_b200.device_path import DevicePath
    from riskaversetrajopt_b200.drone import drone_params as dp
    M = 2_000_000
    dev = torch.device("cuda", 0)
    DWs, masses, obs_Qs = bench.synthetic_drone_samples(M, 5, dev)
    path = DevicePath(_lib.SAA_DRONE, 'saa', S, 0.1, M, device=0)
    path.set_params_drone(dp, dp.OSQP_TOL)
    path.set_samples_drone(masses, DWs, obs_Qs)
    n_rows, n_cols, nnz = path.pattern_sizes()
    assert nnz == 1263 * M + 180 > 2**31 and n_rows == 68 + 61 * M and n_cols == 62 + M
    # int64 column pointers from the library (row indices not materialised: 20 GB on the host)
    indptr = np.empty(n_cols + 1, dtype=np.int64)
    check(lib.saa_pattern_i64(path.handle, 0, indptr.ctypes.data, None), path.handle)
    assert indptr[0] == 0 and indptr[-1] == nnz and np.all(np.diff(indptr) > 0)
    # the int32 entry point must refuse this size
    i32 = np.empty(4, dtype=np.int32)
    assert lib.saa_pattern_i32(path.handle, 0, i32.ctypes.data, i32.ctypes.data) == -1
    for j in range(S - 1):
        for a in range(2):
            assert indptr[j * 3 + a] + 2 == _col(j, a, M)[0]          # 2 final rows precede the sample runs
    assert indptr[60] == 1140 * M + 177                                # y columns start here
    assert indptr[60 + M] - indptr[60] == 62 * M                       # M y columns of 2 + 60 entries
    us = bench.bench_us()
    b = path.assemble(us, 2)
    Ax, u = b['Ax'], b['u']
    assert Ax.numel() == nnz
    row_s0 = 6 + 1 + M
    # samples whose runs straddle element 2^31 in some column, plus the ends and random ones
    idx = [0, 1, 15, 16, M - 1, M - 16, M - 17]
    for j in range(S - 1):
        for a in range(2):
            start, L = _col(j, a, M)
            if start < 2**31 < start + M * L:
                i = (2**31 - start) // L
                idx += [int(i) - 1, int(i), int(i) + 1]
    assert len(idx) > 7                                                # the boundary IS crossed inside runs
    idx = np.unique(np.clip(np.concatenate([idx, np.random.RandomState(1).randint(0, M, 24)]), 0, M - 1))
    it = torch.as_tensor(idx, device=dev)
    ref = DroneOracleB(S, DWs[it].cpu().numpy(), masses[it].cpu().numpy(), obs_Qs[it].cpu().numpy(), 'saa', 0.1)
    _, _, _, g_du, g_up = ref.per_sample(us)
    rel = lambda got, want: float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-12)))
    ub = u[(row_s0 + it[:, None] * 60 + torch.arange(60, device=dev)[None, :])].cpu().numpy()
    assert rel(ub, 0.01 * g_up.reshape(len(idx), -1)) < 1e-9
    for j in range(S - 1):
        for a in range(2):
            start, L = _col(j, a, M)
            pos = start + it[:, None] * L + torch.arange(L, device=dev)[None, :]
            got = Ax[pos].cpu().numpy().reshape(len(idx), 3, L // 3)
            want = 0.01 * g_du[:, :, j + 1:, j * 3 + a]
            assert rel(got, want) < 1e-9, (j, a)
    # constants beyond 2^31: t column (last) and the y columns
    tcol = int(indptr[60 + M + 1])
    assert Ax[tcol].item() == M * 0.1 and bool((Ax[tcol + 1:tcol + 1 + 1000] == -0.01).all())
    ycol = int(indptr[60 + M - 1])
    assert Ax[ycol].item() == 1.0 and Ax[ycol + 1].item() == -1.0 and bool((Ax[ycol + 2:ycol + 62] == -0.01).all())
