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
