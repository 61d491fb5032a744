"""Host QP stand-in (context, not the accelerated path): correctness on small problems and on
the oracle-built car SCP."""
import numpy as np
import scipy.sparse as sp

from riskaversetrajopt_b200.qp import make_solver


def test_small_qp_matches_known_solution():
    P = sp.csc_matrix(np.array([[4., 1], [1, 2]])); q = np.array([1., 1])
    A = sp.csc_matrix(np.array([[1., 1], [1, 0], [0, 1]]))
    l, u = np.array([1., 0, 0]), np.array([1., 0.7, 0.7])
    for name in ('admm', 'highs'):
        s = make_solver(name)
        s.setup(P, q, A, l, u, eps_abs=1e-7, eps_rel=1e-7)
        r = s.solve()
        assert r.info.status == 'solved' and np.allclose(r.x, [0.3, 0.7], atol=1e-5)


def test_update_interface_and_warm_start():
    rs = np.random.RandomState(0)
    n, m = 8, 12
    Q = rs.randn(n, n); P = sp.csc_matrix(Q @ Q.T + np.eye(n)); q = rs.randn(n)
    A = sp.csc_matrix(rs.randn(m, n)); l, u = -np.ones(m), np.ones(m)
    s = make_solver('admm'); s.setup(P, q, A, l, u, eps_abs=1e-8, eps_rel=1e-8)
    x0 = s.solve().x
    A2 = A.copy(); A2.data *= 1.1
    s.update(l=-2 * np.ones(m), u=2 * np.ones(m)); s.update(Ax=A2.data)
    x1 = s.solve().x
    s2 = make_solver('admm'); s2.setup(P, q, A2, -2 * np.ones(m), 2 * np.ones(m), eps_abs=1e-8, eps_rel=1e-8)
    assert np.allclose(x1, s2.solve().x, atol=1e-5) and not np.allclose(x0, x1)


def test_car_scp_converges_on_oracle_matrices():
    """The reference's SCP loop (car/driving.py:482-513) on oracle-built matrices."""
    from oracle.oracle_b import CarOracleB
    from riskaversetrajopt_b200.car import driving_params as cp
    from riskaversetrajopt_b200.car.driving import sample_uncertain_parameters
    st = np.random.get_state(); np.random.seed(0)
    smp = sample_uncertain_parameters(20, 'saa'); np.random.set_state(st)
    b = CarOracleB(*smp, 'saa', 0.05)
    n = 42 + 20
    P = sp.lil_matrix((n, n))
    for t in range(20):
        P[2 * t, 2 * t] = 2 * cp.dt; P[2 * t + 1, 2 * t + 1] = 2 * cp.dt / 3
    P[-2, -2] = 1000.; q = np.zeros(n); q[-2] = 1000.
    us = np.full((20, 2), 0.01)
    for it in range(10):
        A, l, u = b.get_constraints_coeffs(us, it)
        l = np.where(np.isnan(l), -np.inf, l)
        if it < 2:
            sol = make_solver('admm'); sol.setup(sp.csc_matrix(P), q, A, l, u, eps_abs=3e-4, eps_rel=3e-4, polish=True)
        else:
            sol.update(l=l, u=u); sol.update(Ax=A.data)
        r = sol.solve()
        new = np.reshape(r.x[:40], (2, 20), 'F').T
        err = np.mean(np.linalg.norm(new - us, axis=-1)) / np.mean(np.linalg.norm(new, axis=-1))
        us = new
    assert r.info.status == 'solved' and err < 1e-2 and r.x[-1] <= 1e-6
    assert np.allclose(b.rollout(us)[0, -1, :4], [20, 0.1, 4.1, 0], atol=1e-3)


def test_kkt_polish_gives_the_exact_solution_when_the_active_set_is_right():
    """polish='kkt' (OSQP's polishing step): coarse ADMM + active-set KKT solve with iterative refinement."""
    rs = np.random.RandomState(3)
    n, m = 10, 30
    Q = rs.randn(n, n); P = sp.csc_matrix(Q @ Q.T + np.eye(n)); q = rs.randn(n)
    A = sp.csc_matrix(rs.randn(m, n)); l, u = -0.3 * np.ones(m), 0.3 * np.ones(m)
    ref = make_solver('admm'); ref.setup(P, q, A, l, u, eps_abs=1e-10, eps_rel=1e-10)
    x_ref = ref.solve().x
    s = make_solver('admm'); s.setup(P, q, A, l, u, eps_abs=1e-3, eps_rel=1e-3, polish='kkt')
    r = s.solve()
    assert r.info.status == 'solved' and r.info.polished
    assert np.max(np.abs(r.x - x_ref)) < 1e-8
    coarse = make_solver('admm'); coarse.setup(P, q, A, l, u, eps_abs=1e-3, eps_rel=1e-3)
    assert np.max(np.abs(coarse.solve().x - x_ref)) > 1e-6          # what polishing bought


def test_infeasibility_certificates():
    """Infeasible and unbounded programs end with OSQP's status strings instead of running to max_iter."""
    P = sp.csc_matrix(np.eye(1)); q = np.zeros(1)
    A = sp.csc_matrix(np.array([[1.0], [1.0]]))
    s = make_solver('admm'); s.setup(P, q, A, np.array([1.0, -np.inf]), np.array([np.inf, 0.0]), max_iter=5000)
    r = s.solve()
    assert r.info.status == 'primal infeasible' and r.info.iter < 5000
    P0 = sp.csc_matrix((2, 2)); q0 = np.array([-1.0, 0.0])
    A0 = sp.csc_matrix(np.eye(2))
    s = make_solver('admm'); s.setup(P0, q0, A0, np.zeros(2), np.array([np.inf, 1.0]), max_iter=5000)
    r = s.solve()
    assert r.info.status == 'dual infeasible' and r.info.iter < 5000
