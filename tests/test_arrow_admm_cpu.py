"""The structure-exploiting ADMM (oracle/arrow_admm.py: Sherman-Morrison for the CVaR row + Schur complement on
the dense variables) is the OSQP iteration of qp.OSQPLike with a different linear solve: same scaling, same
iterates, same iteration counts.  (The CUDA implementation is checked against both in tests/test_gpu_qp.py.)"""
import numpy as np
import scipy.sparse as sp


def _drone_problem(M):
    from oracle.oracle_b import DroneOracleB
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)
    ref = DroneOracleB(dp.S, DWs, masses, obs_Qs, 'saa', 0.1)
    n = 62 + M
    P = sp.lil_matrix((n, n))
    P[:60, :60] = np.kron(np.eye(20), 2 * dp.dt * np.asarray(dp.R))
    P[n - 2, n - 2] = 1e4
    q = np.zeros(n); q[-2] = 1e4
    return ref, sp.csc_matrix(P), q


def test_structured_solve_reproduces_osqplike_iterates():
    from oracle.arrow_admm import ArrowQP, ArrowADMM
    from riskaversetrajopt_b200.qp import OSQPLike
    M = 24
    ref, P, q = _drone_problem(M)
    us = np.tile(np.array([0.01, 0.01, 0.0]), (20, 1))
    A, l, u = ref.get_constraints_coeffs(us, 2)
    host = OSQPLike().setup(P, q, A, l, u, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
    arrow = ArrowADMM(ArrowQP(P, q, A, l, u, 60, M, 6, 60), eps_abs=1e-4, eps_rel=1e-4)
    E = np.concatenate([arrow.EF, [arrow.Ec], arrow.Ey, arrow.Es.ravel(), [arrow.Esl], arrow.Ect])
    assert np.allclose(E, host.E, rtol=1e-12) and np.allclose(arrow.Du, host.D[:60], rtol=1e-12)
    assert np.allclose(arrow.Dy, host.D[60:60 + M], rtol=1e-12) and np.isclose(arrow.c, host.c, rtol=1e-12)
    for it in range(4):
        A, l, u = ref.get_constraints_coeffs(us, it)
        host.update(l=l, u=u); host.update(Ax=A.data)
        arrow.update(A=A, l=l, u=u)
        r0, r1 = host.solve(), arrow.solve()
        assert r0.info.status == r1.info.status == 'solved'
        assert r0.info.iter == r1.info.iter, (it, r0.info.iter, r1.info.iter)
        assert np.max(np.abs(r0.x - r1.x)) < 1e-6 * max(1.0, np.max(np.abs(r0.x)))
        us = np.reshape(r0.x[:60], (3, 20), 'F').T


def test_structured_solve_on_the_car_program():
    """Same check on the car's CVaR program (one separation row per step: R = 20, nu = 40, 4 final rows)."""
    from oracle.arrow_admm import ArrowQP, ArrowADMM
    from oracle.oracle_b import CarOracleB
    from riskaversetrajopt_b200.car import driving_params as cp
    from riskaversetrajopt_b200.car.driving import sample_uncertain_parameters
    from riskaversetrajopt_b200.qp import OSQPLike
    M = 16
    st = np.random.get_state(); np.random.seed(0)
    smp = sample_uncertain_parameters(M, 'saa'); np.random.set_state(st)
    ref = CarOracleB(*smp, 'saa', 0.1)
    n = 42 + M
    P = sp.lil_matrix((n, n))
    for t in range(20):
        P[2 * t, 2 * t] = 2 * cp.dt; P[2 * t + 1, 2 * t + 1] = 2 * cp.dt / 3
    P[n - 2, n - 2] = 1000.0
    P = sp.csc_matrix(P)
    q = np.zeros(n); q[-2] = 1000.0
    us = np.full((20, 2), 0.01) + 0.05 * np.random.RandomState(1).randn(20, 2)
    A, l, u = ref.get_constraints_coeffs(us, 1)
    host = OSQPLike().setup(P, q, A, l, u, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
    arrow = ArrowADMM(ArrowQP(P, q, A, l, u, 40, M, 4, 20), eps_abs=1e-4, eps_rel=1e-4)
    for it in (1, 2, 3):
        A, l, u = ref.get_constraints_coeffs(us, it)
        host.update(l=l, u=u); host.update(Ax=A.data)
        arrow.update(A=A, l=l, u=u)
        r0, r1 = host.solve(), arrow.solve()
        assert r0.info.status == r1.info.status == 'solved'
        assert r0.info.iter == r1.info.iter, (it, r0.info.iter, r1.info.iter)
        assert np.max(np.abs(r0.x - r1.x)) < 1e-6 * max(1.0, np.max(np.abs(r0.x)))
        us = np.reshape(r0.x[:40], (2, 20), 'F').T
