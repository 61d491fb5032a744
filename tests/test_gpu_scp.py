"""SCP convergence parity: the same host QP solver driven by GPU-built and by oracle-built
(P, q, A, l, u) must reach the same trajectory (north_star: "the same converged SCP trajectory")."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _scp(get_coeffs, P, q, us0, iters, reshape):
    from riskaversetrajopt_b200.qp import make_solver
    s = make_solver('admm')
    A, l, u = get_coeffs(us0, 2)
    s.setup(P, q, A, l, u, eps_abs=1e-5, eps_rel=1e-5, warm_start=True)
    us, hist = us0, []
    for it in range(iters):
        A, l, u = get_coeffs(us, it)
        s.update(l=l, u=u)
        s.update(Ax=A.data)
        r = s.solve()
        assert r.info.status == 'solved'
        us = reshape(r.x)
        hist.append((us, r.x[-1]))
    return hist


def test_drone_scp_same_trajectory(drone_seed0):
    from oracle.oracle_b import DroneOracleB
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_risk import Model
    DWs, masses, obs_Qs = drone_seed0
    model = Model(dp.S, DWs, masses, obs_Qs, 'saa', 0.1)
    ref = DroneOracleB(dp.S, DWs, masses, obs_Qs, 'saa', 0.1)
    P, q = model.get_objective_coeffs()
    reshape = lambda x: np.reshape(x[:60], (3, 20), 'F').T
    us0 = model.initial_guess_us_mat()
    hg = _scp(model.get_constraints_coeffs, P, q, us0, 8, reshape)
    hr = _scp(ref.get_constraints_coeffs, P, q, us0, 8, reshape)
    for (ug, tg), (ur, tr) in zip(hg, hr):
        assert np.allclose(ug, ur, rtol=1e-6, atol=1e-7) and abs(tg - tr) < 1e-6
    # and the SCP actually moved away from the initial guess towards the goal
    Xs = model.us_to_state_trajectories(hg[-1][0])
    assert np.linalg.norm(Xs[:, -1, :2].mean(axis=0)) < 0.2


def test_model_scp_glue_runs(drone_seed0):
    """define_problem / update_problem / solve as the reference's driver calls them (:506-531)."""
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_risk import Model, L2_error_us
    DWs, masses, obs_Qs = (x[:20] for x in drone_seed0)
    model = Model(dp.S, DWs, masses, obs_Qs, 'saa', 0.2)
    us_prev = model.initial_guess_us_mat()
    model.define_problem(us_prev)
    for it in range(5):
        model.update_problem(us_prev, it)
        us, t_risk = model.solve(verbose=False)
        err = L2_error_us(us, us_prev)
        us_prev = us
    assert us.shape == (20, 3) and np.isfinite(err) and model.res.info.status == 'solved'


def test_car_scp_glue_runs():
    from riskaversetrajopt_b200.car.driving import Model
    np.random.seed(0)
    model = Model(20, 'saa', 0.1)
    us_prev = model.initial_guess_us_mat()
    for it in range(4):
        model.define_problem(us_prev, it)
        us, t_risk = model.solve()
        us_prev = us
    assert us.shape == (20, 2) and np.all(np.isfinite(us))


def test_tail_subproblem_as_the_update_problem_path():
    """Model.define_problem(..., tail=...) feeds the host QP the tail-reduced subproblem; at convergence the
    left-out margin is <= 0 (the reduction is exact there) and the SCP reaches the full problem's fixed
    point.  With max_resolves > 0 a positive margin doubles K and re-solves at the same iterate."""
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_risk import Model
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=120)
    full, red = (Model(dp.S, DWs, masses, obs_Qs, 'saa', 0.1) for _ in range(2))
    us_f = us_r = full.initial_guess_us_mat()
    full.define_problem(us_f, tail=False)
    red.define_problem(us_r, tail=dict(margin=1.0, max_resolves=1))          # K = 2 alpha M
    assert red._tail.tail.K == 24 and full._tail is None
    margins = []
    for it in range(10):
        full.update_problem(us_f, it); us_f, t_f = full.solve(verbose=False)
        red.update_problem(us_r, it); us_r, t_r = red.solve(verbose=False)
        margins.append(red.left_out_margin)
    assert margins[0] == -np.inf and margins[1] == -np.inf           # relaxed iterations: nothing to check
    assert red._tail.resolves >= 1 and red._tail.tail.K > 24          # early iterates: every sample violates -> K grew
    assert margins[-1] <= 0                                           # at convergence the reduction is exact ...
    assert np.max(np.abs(us_r - us_f)) < 5e-3 and abs(t_r - t_f) < 5e-3   # ... same fixed point (OSQP_TOL = 1e-3)
