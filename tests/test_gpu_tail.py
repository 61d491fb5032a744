"""Tail-reduced subproblem: device selection of the K largest Z_i, the
means-only pass, and the reduced matrix = the full matrix with the other samples deleted."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _select_ref(Z, K):
    """K largest, ties towards the smaller index, result ascending."""
    order = np.lexsort((np.arange(len(Z)), -Z.astype(np.float64)))
    return np.sort(order[:K])


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_select_tail_matches_sort(precision):
    import torch
    from riskaversetrajopt_b200 import _lib
    from riskaversetrajopt_b200._lib import lib, check
    from riskaversetrajopt_b200.device_path import DevicePath
    rs = np.random.RandomState(0)
    dt = torch.float64 if precision == "fp64" else torch.float32
    cases = []
    for n in (1, 31, 4096, 4097, 100003):
        cases.append(rs.randn(n))                                   # generic
        cases.append(np.round(rs.randn(n) * 3))                     # many ties, signed zeros
        z = rs.randn(n); z[rs.rand(n) < 0.3] = -np.inf; z[rs.rand(n) < 0.01] = np.inf
        cases.append(z)
        cases.append(np.full(n, -0.25))                             # all equal
        cases.append(rs.randn(n) * 1e-300 if precision == "fp64" else rs.randn(n) * 1e-38)   # tiny / denormal
    for z in cases:
        n = len(z)
        path = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, n, precision=precision)
        Z = torch.as_tensor(z).to('cuda', dtype=dt)
        zc = Z.cpu().numpy()
        for K in sorted({1, min(n, 2), max(1, n // 10), max(1, n // 2), n}):
            idx = torch.full((K,), -1, dtype=torch.int64, device='cuda')
            check(lib.saa_select_tail(path._h, Z.data_ptr(), K, idx.data_ptr(), path._stream()), path._h)
            got = idx.cpu().numpy()
            assert np.array_equal(got, _select_ref(zc, K)), (n, K)
        path.close()
    # argument checks
    path = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, 8, precision=precision)
    Z = torch.zeros(8, dtype=dt, device='cuda'); idx = torch.zeros(9, dtype=torch.int64, device='cuda')
    for K in (0, 9):
        with pytest.raises(_lib.SaaError):
            check(lib.saa_select_tail(path._h, Z.data_ptr(), K, idx.data_ptr(), path._stream()), path._h)


def _submatrix(A, l, u, idx, M, R, nu, n_fin):
    """Full (A, l, u) restricted to the samples idx: rows [final | cvar | -y_i | sample rows | slack |
    controls], columns [u | y | slack | t]  (reference drone_risk.py:327-368)."""
    idx = np.asarray(idx)
    rows = np.concatenate([np.arange(n_fin), [n_fin], n_fin + 1 + idx,
                           (n_fin + 1 + M + idx[:, None] * R + np.arange(R)[None, :]).ravel(),
                           [n_fin + 1 + M + M * R], n_fin + 2 + M + M * R + np.arange(nu)])
    cols = np.concatenate([np.arange(nu), nu + idx, [nu + M, nu + M + 1]])
    As = A.tocsr()[rows][:, cols].tocsc()
    As.sort_indices()
    return As, l[rows], u[rows]


@pytest.mark.parametrize("scp_iter", [0, 2])
def test_drone_reduced_matrix_is_the_full_one_restricted(scp_iter):
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    from riskaversetrajopt_b200.drone.drone_risk import Model
    from oracle.oracle_b import DroneOracleB
    np.random.seed(3)
    M = 333
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)
    model = Model(dp.S, DWs, masses, obs_Qs, 'saa', 0.1)
    us = model.initial_guess_us_mat() + 0.05 * np.random.RandomState(1).randn(dp.S, dp.n_u)
    A, l, u = model.get_constraints_coeffs(us, scp_iter)
    tail = model.tail_subproblem(margin=0.5)
    assert tail.K == 50
    Ar, lr, ur, idx = tail.get_constraints_coeffs(us, scp_iter)
    # the selection: K largest max-constraint values of the oracle's rollout
    _, Zmax = model.monte_carlo_constraints(us)
    assert np.array_equal(idx, _select_ref(Zmax, tail.K))
    As, ls, us_ = _submatrix(A, l, u, idx, M, 60, 60, 6)
    assert Ar.shape == As.shape and np.array_equal(Ar.indptr, As.indptr) and np.array_equal(Ar.indices, As.indices)
    fin = Ar.indices < 6                       # expectation rows: other summation order
    assert np.array_equal(Ar.data[~fin], As.data[~fin])          # bitwise: same instructions, same inputs
    assert np.allclose(Ar.data[fin], As.data[fin], rtol=1e-12, atol=1e-15)
    assert np.array_equal(lr[6:], ls[6:]) and np.array_equal(ur[6:], us_[6:])
    assert np.allclose(lr[:6], ls[:6], rtol=1e-12, atol=1e-15) and np.allclose(ur[:6], us_[:6], rtol=1e-12, atol=1e-15)
    # ... and against the oracle
    ref = DroneOracleB(dp.S, DWs, masses, obs_Qs, 'saa', 0.1)
    Ao, lo, uo = ref.get_constraints_coeffs(us, scp_iter)
    Aos, los, uos = _submatrix(Ao, lo, uo, idx, M, 60, 60, 6)
    scale = np.maximum(np.abs(Aos.data), 1e-12)
    assert np.max(np.abs(Ar.data - Aos.data) / scale) < 1e-9
    assert np.allclose(lr, los, rtol=1e-9, atol=1e-12) and np.allclose(ur, uos, rtol=1e-9, atol=1e-12)


def test_car_reduced_matrix_is_the_full_one_restricted():
    from riskaversetrajopt_b200.car.driving import Model
    np.random.seed(0)
    M = 120
    model = Model(M, 'saa', 0.1)
    us = model.initial_guess_us_mat() + 0.05 * np.random.RandomState(2).randn(20, 2)
    A, l, u = model.get_constraints_coeffs(us, 2)
    tail = model.tail_subproblem(K=30)
    Ar, lr, ur, idx = tail.get_constraints_coeffs(us, 2)
    _, Zmax = model.monte_carlo_constraints(us)
    assert np.array_equal(idx, _select_ref(Zmax, 30))
    As, ls, us_ = _submatrix(A, l, u, idx, M, 20, 40, 4)
    assert Ar.shape == As.shape and np.array_equal(Ar.indptr, As.indptr) and np.array_equal(Ar.indices, As.indices)
    assert np.array_equal(Ar.data, As.data)
    assert np.array_equal(lr, ls) and np.array_equal(ur, us_)


def test_means_only_pass_matches_assemble():
    import torch
    from riskaversetrajopt_b200._lib import lib, check
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    from riskaversetrajopt_b200.drone.drone_risk import Model
    np.random.seed(5)
    M = 1000
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)
    model = Model(dp.S, DWs, masses, obs_Qs, 'saa', 0.1)
    us = model.initial_guess_us_mat() + 0.1 * np.random.RandomState(7).randn(dp.S, dp.n_u)
    p = model.path
    p.assemble(us, 2, finalize=False)
    ref = p.mean_sums.clone()
    p.mean_sums.zero_()
    Z = torch.empty(M, dtype=torch.float64, device=p.device)
    check(lib.saa_linearize_means(p._h, np.ascontiguousarray(us).ctypes.data, Z.data_ptr(), p.mean_sums.data_ptr(),
                                  p._stream()), p._h)
    assert torch.allclose(p.mean_sums, ref, rtol=1e-12, atol=1e-12)
    _, Zmax = model.monte_carlo_constraints(us)
    assert np.allclose(Z.cpu().numpy(), Zmax + dp.OSQP_TOL, rtol=0, atol=1e-15)


def test_reduced_qp_has_the_full_qp_solution(drone_seed0):
    """With the active tail inside the selection the two QPs have the same minimiser."""
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_risk import Model
    from riskaversetrajopt_b200.qp import make_solver
    DWs, masses, obs_Qs = drone_seed0
    M = len(masses)
    model = Model(dp.S, DWs, masses, obs_Qs, 'saa', 0.2)
    us = model.initial_guess_us_mat()
    n_it = 9
    for it in range(n_it):                        # full SCP steps until the iterates have settled
        P, q = model.get_objective_coeffs()
        A, l, u = model.get_constraints_coeffs(us, it)
        s = make_solver('admm'); s.setup(P, q, A, l, u, eps_abs=1e-7, eps_rel=1e-7)
        xf = s.solve().x
        if it == n_it - 1:
            break
        us = np.reshape(xf[:60], (3, 20), 'F').T
    tail = model.tail_subproblem(K=M // 2)
    Ar, lr, ur, idx = tail.get_constraints_coeffs(us, n_it - 1)
    # premise: the samples left out are inactive at the full QP's solution (their y is 0)
    out = np.setdiff1d(np.arange(M), idx)
    assert np.all(np.abs(xf[60 + out]) < 2e-5)
    Pr, qr = tail.get_objective_coeffs(P, q)
    s = make_solver('admm'); s.setup(Pr, qr, Ar, lr, ur, eps_abs=1e-7, eps_rel=1e-7)
    xr = s.solve().x
    assert np.allclose(xr[:60], xf[:60], atol=2e-5) and abs(xr[-1] - xf[-1]) < 2e-5
    assert np.allclose(xr[60:60 + tail.K], xf[60 + idx], atol=2e-5)
    # ... which the a-posteriori check of the reduced solve confirms on the device
    assert tail.left_out_margin(xr[-1]) <= 0.0


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_tail_edge_sizes_and_fp32(precision):
    """K = M reproduces the full matrix; K = 1 keeps the worst sample; FP32 handles select on FP32 keys."""
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    from riskaversetrajopt_b200.drone.drone_risk import Model
    np.random.seed(9)
    M = 77
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)
    model = Model(dp.S, DWs, masses, obs_Qs, 'saa', 0.1, precision=precision)
    us = model.initial_guess_us_mat() + 0.05 * np.random.RandomState(3).randn(dp.S, dp.n_u)
    A, l, u = model.get_constraints_coeffs(us, 2)
    _, Zmax = model.monte_carlo_constraints(us)
    tol = dict(rtol=1e-12, atol=1e-15) if precision == "fp64" else dict(rtol=1e-5, atol=1e-7)
    for K in (M, 1, 16, 17):
        tail = model.tail_subproblem(K=K)
        Ar, lr, ur, idx = tail.get_constraints_coeffs(us, 2)
        assert np.array_equal(idx, _select_ref(Zmax, K))
        As, ls, us_ = _submatrix(A, l, u, idx, M, 60, 60, 6)
        assert np.array_equal(Ar.indptr, As.indptr) and np.array_equal(Ar.indices, As.indices)
        fin = Ar.indices < 6
        assert np.array_equal(Ar.data[~fin], As.data[~fin])
        assert np.allclose(Ar.data[fin], As.data[fin], **tol)
        assert np.array_equal(lr[6:], ls[6:]) and np.array_equal(ur[6:], us_[6:])
    with pytest.raises(ValueError):
        model.tail_subproblem(K=M + 1)
    base = Model(dp.S, DWs, masses, obs_Qs, 'baseline', 0.1)
    with pytest.raises(ValueError):
        base.tail_subproblem()


def _select_over_shards(paths, Zs, K_total):
    """dist.exact_global_select with the ranks played by several handles on one GPU (the all-reduce is a
    tensor sum, the all-gather a stack)."""
    import torch
    from riskaversetrajopt_b200._lib import lib, check
    hists = [torch.zeros(256, dtype=torch.int32, device='cuda') for _ in paths]
    for p in paths:
        check(lib.saa_select_begin(p._h, K_total, p._stream()), p._h)
    for ps in range(int(lib.saa_select_passes(paths[0]._h))):
        for p, Z, h in zip(paths, Zs, hists):
            check(lib.saa_select_pass_hist(p._h, Z.data_ptr(), ps, h.data_ptr(), p._stream()), p._h)
        tot = torch.stack(hists).sum(0).to(torch.int32)
        for p, h in zip(paths, hists):
            h.copy_(tot)
            check(lib.saa_select_pass_pick(p._h, h.data_ptr(), ps, p._stream()), p._h)
    cnts = []
    for p, Z in zip(paths, Zs):
        c = torch.zeros(2, dtype=torch.int64, device='cuda')
        check(lib.saa_select_counts(p._h, Z.data_ptr(), c.data_ptr(), p._stream()), p._h)
        cnts.append(c.cpu().numpy())
    rem = K_total - sum(int(c[0]) for c in cnts)
    out = []
    for p, Z, c in zip(paths, Zs, cnts):
        take = int(min(c[1], max(rem, 0))); rem -= take
        idx = torch.full((max(1, int(c[0]) + take),), -1, dtype=torch.int64, device='cuda')
        check(lib.saa_select_finish(p._h, Z.data_ptr(), take, idx.data_ptr(), p._stream()), p._h)
        out.append(idx[:int(c[0]) + take].cpu().numpy())
    return out


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_exact_selection_over_shards(precision):
    """The histogram-exchanging radix select picks the global top K whatever the split of the values over
    the shards (ties to the lower global index), and with one shard it is saa_select_tail."""
    import torch
    from riskaversetrajopt_b200 import _lib
    from riskaversetrajopt_b200.device_path import DevicePath
    rs = np.random.RandomState(1)
    dt = torch.float64 if precision == "fp64" else torch.float32
    for n, splits in ((1000, (1000,)), (1000, (500, 500)), (1003, (1, 1002)), (5000, (1200, 1300, 2500)), (64, (32, 32))):
        for kind in ("generic", "ties", "skew", "equal"):
            z = rs.randn(n)
            if kind == "ties":
                z = np.round(z * 2)
            elif kind == "skew":
                z = np.sort(z)[::-1].copy()                          # the whole tail on the first shard
            elif kind == "equal":
                z[:] = 0.5
            zc = torch.as_tensor(z).to(dt).numpy()
            bounds = np.concatenate([[0], np.cumsum(splits)])
            paths = [DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, c, precision=precision) for c in splits]
            Zs = [torch.as_tensor(zc[a:b]).cuda() for a, b in zip(bounds[:-1], bounds[1:])]
            for K in (1, n // 10, n // 2, n):
                parts = _select_over_shards(paths, Zs, K)
                got = np.concatenate([p + a for p, a in zip(parts, bounds[:-1])])
                assert np.array_equal(got, _select_ref(zc, K)), (n, splits, kind, K)
            for p in paths:
                p.close()


def test_active_subset_of_a_capacity_handle(drone_seed0):
    """saa_set_active: a handle created for K samples assembles K' <= K of them at any offset of a K_out
    matrix; K' = 0 is a no-op (the rank contributes nothing to the tail)."""
    import torch
    from riskaversetrajopt_b200 import _lib
    from riskaversetrajopt_b200._lib import lib, check
    from riskaversetrajopt_b200.device_path import DevicePath
    from riskaversetrajopt_b200.drone import drone_params as dp
    DWs, masses, obs_Qs = (x[:40] for x in drone_seed0)
    us = 0.1 * np.random.RandomState(0).randn(20, 3)
    full = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, 40)
    full.set_params_drone(dp, dp.OSQP_TOL); full.set_samples_drone(masses, DWs, obs_Qs)
    sub = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, 32, M_global=40)
    sub.set_params_drone(dp, dp.OSQP_TOL)
    with pytest.raises(_lib.SaaError):
        sub.set_active(33, 40, 0)
    for pick, first in (([3, 5, 8, 13, 21], 7), (list(range(32)), 0), ([39], 11), ([], 4)):
        K = len(pick)
        idx = torch.as_tensor(pick + [0], dtype=torch.int64, device='cuda')
        sub.set_active(K, 12 + K, first)
        check(lib.saa_gather_samples(sub._h, full._h, idx.data_ptr(), sub._stream()), sub._h)
        n_rows, _, nnz = sub.pattern_sizes(False)
        out = dict(Ax=torch.full((nnz,), 7.0, dtype=torch.float64, device='cuda'),
                   l=torch.full((n_rows,), 7.0, dtype=torch.float64, device='cuda'),
                   u=torch.full((n_rows,), 7.0, dtype=torch.float64, device='cuda'), const_state=None)
        sub.assemble(us, 2, out=out, write_shared=False, finalize=False)
        # the same samples through a handle created for exactly K samples
        if K == 0:
            assert all(bool((out[k] == 7.0).all()) for k in ('Ax', 'l', 'u'))
            continue
        ref = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, K, M_global=40)
        ref.set_params_drone(dp, dp.OSQP_TOL)
        ref.set_samples_drone(masses[pick], DWs[pick], obs_Qs[pick])
        ref.set_output_geometry(12 + K, first)
        want = dict(Ax=torch.full((nnz,), 7.0, dtype=torch.float64, device='cuda'),
                    l=torch.full((n_rows,), 7.0, dtype=torch.float64, device='cuda'),
                    u=torch.full((n_rows,), 7.0, dtype=torch.float64, device='cuda'), const_state=None)
        ref.assemble(us, 2, out=want, write_shared=False, finalize=False)
        for k in ('Ax', 'l', 'u'):
            assert torch.equal(out[k], want[k]), (k, K, first)
        ref.close()
    sub.close(); full.close()
