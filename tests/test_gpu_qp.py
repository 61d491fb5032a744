"""Device-resident ADMM (riskaversetrajopt_b200/device_qp.py, csrc/qp_kernels.cuh) against the host solver
it restates (qp.OSQPLike, same algorithm with a sparse factorisation) on the matrices the GPU assembled:
same Ruiz scaling, same iteration counts, same solution -- and the same SCP trajectory."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _drone_model(M, alpha=0.1):
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_risk import Model
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)
    return Model(dp.S, DWs, masses, obs_Qs, 'saa', alpha)


@pytest.mark.parametrize("M", [1, 37, 200])
def test_device_qp_matches_host_admm_drone(M):
    from riskaversetrajopt_b200.device_qp import DeviceQP
    from riskaversetrajopt_b200.qp import OSQPLike
    model = _drone_model(M)
    P, q = model.get_objective_coeffs()
    us = model.initial_guess_us_mat()
    A, l, u = model.get_constraints_coeffs(us, 2)
    host = OSQPLike().setup(P, q, A, l, u, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
    dq = DeviceQP(model.path, eps_abs=1e-4, eps_rel=1e-4).setup(P, q, model.path.assemble(us, 2))
    # the Ruiz scaling is the same sequence of operations
    nu = 60
    assert np.allclose(dq.Du, host.D[:nu], rtol=1e-12) and np.isclose(dq.c, host.c, rtol=1e-12)
    assert np.allclose(dq.st['Dy'].cpu().numpy(), host.D[nu:nu + M], rtol=1e-12)
    assert np.allclose(dq.st['Es'].cpu().numpy(), host.E[7 + M:7 + M + 60 * M], rtol=1e-12)
    assert np.allclose(dq.st['Ey'].cpu().numpy(), host.E[7:7 + M], rtol=1e-12)
    first_it = 0
    for it in range(5):
        A, l, u = model.get_constraints_coeffs(us, it)
        host.update(l=l, u=u); host.update(Ax=A.data)
        dq.update(model.path.assemble(us, it))
        r0, r1 = host.solve(), dq.solve()
        assert r0.info.status == r1.info.status == 'solved', (it, r0.info.status, r1.info.status)
        assert abs(r0.info.iter - r1.info.iter) <= 10, (it, r0.info.iter, r1.info.iter)   # one termination check
        # same iterates up to rounding at first; after hundreds of warm-started ADMM iterations the two linear
        # solves' rounding differences grow, but stay well inside the solver tolerance (1e-4)
        assert np.max(np.abs(r0.x - r1.x)) < (1e-7 if it == first_it else 1e-4), it
        us = model.convert_us_vec_to_us_mat(r0.x[:nu])


def test_device_qp_matches_host_admm_car():
    from riskaversetrajopt_b200.car.driving import Model
    from riskaversetrajopt_b200.device_qp import DeviceQP
    from riskaversetrajopt_b200.qp import OSQPLike
    np.random.seed(0)
    model = Model(40, 'saa', 0.1)
    P, q = model.get_objective_coeffs()
    us = model.initial_guess_us_mat() + 0.05 * np.random.RandomState(1).randn(20, 2)
    A, l, u = model.get_constraints_coeffs(us, 1)
    host = OSQPLike().setup(P, q, A, l, u, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
    dq = DeviceQP(model.path, eps_abs=1e-4, eps_rel=1e-4).setup(P, q, model.path.assemble(us, 1))
    first_it = 1
    for it in range(1, 5):
        A, l, u = model.get_constraints_coeffs(us, it)
        host.update(l=l, u=u); host.update(Ax=A.data)
        dq.update(model.path.assemble(us, it))
        r0, r1 = host.solve(), dq.solve()
        assert r0.info.status == r1.info.status == 'solved'
        assert abs(r0.info.iter - r1.info.iter) <= 10, (it, r0.info.iter, r1.info.iter)   # one termination check
        # same iterates up to rounding at first; after hundreds of warm-started ADMM iterations the two linear
        # solves' rounding differences grow, but stay well inside the solver tolerance (1e-4)
        assert np.max(np.abs(r0.x - r1.x)) < (1e-7 if it == first_it else 1e-4), it
        us = model.convert_us_vec_to_us_mat(r0.x[:40])


def test_scp_with_the_device_solver_reaches_the_host_solvers_trajectory():
    """Model.define_problem(solver='device'): the SCP never brings the matrix to the host."""
    from riskaversetrajopt_b200.drone.drone_risk import L2_error_us
    host, dev = _drone_model(64), _drone_model(64)
    us_h = us_d = host.initial_guess_us_mat()
    opts = dict(eps_abs=1e-6, eps_rel=1e-6, polish=False, max_iter=200000)       # the same ADMM on both sides
    host.define_problem(us_h, tail=False, solver_opts=opts)
    dev.define_problem(us_d, solver='device', solver_opts=opts)
    for it in range(8):
        host.update_problem(us_h, it); us_h, t_h = host.solve(verbose=False)
        dev.update_problem(us_d, it); us_d, t_d = dev.solve(verbose=False)
        assert dev.res.info.status == 'solved'
        assert np.max(np.abs(us_h - us_d)) < 1e-4 and abs(t_h - t_d) < 1e-4, it      # rounding differences grow along the SCP
    assert L2_error_us(us_d, us_h) < 1e-5


def test_tail_subproblem_solved_on_the_device():
    """define_problem(tail=..., solver='device'): selection, assembly AND the solve of the tail-reduced QP stay on
    the device; the SCP follows the host-solved tail path."""
    host, dev = _drone_model(600), _drone_model(600)
    us_h = us_d = host.initial_guess_us_mat()
    opts = dict(eps_abs=1e-6, eps_rel=1e-6, polish=False, max_iter=200000)       # the same ADMM on both sides
    host.define_problem(us_h, tail=0.5, solver_opts=opts)
    dev.define_problem(us_d, tail=0.5, solver='device', solver_opts=opts)
    assert dev._tail.tail.K == host._tail.tail.K
    for it in range(6):
        host.update_problem(us_h, it); us_h, t_h = host.solve(verbose=False)
        dev.update_problem(us_d, it); us_d, t_d = dev.solve(verbose=False)
        assert dev.res.info.status == 'solved' and host.res.info.status == 'solved'
        assert np.max(np.abs(us_h - us_d)) < 1e-4 and abs(t_h - t_d) < 1e-4, it
    assert dev.left_out_margin == host.left_out_margin or abs(dev.left_out_margin - host.left_out_margin) < 1e-3


def test_device_qp_other_horizon():
    """S = 12 goes through the run-time-horizon QP kernels (the compile-time ones are for the reference's S = 20)."""
    from riskaversetrajopt_b200.device_qp import DeviceQP
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_risk import Model
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    from riskaversetrajopt_b200.qp import OSQPLike
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=30)
    model = Model(12, DWs[:, :12], masses, obs_Qs, 'saa', 0.1)
    P, q = model.get_objective_coeffs()
    us = model.initial_guess_us_mat()
    A, l, u = model.get_constraints_coeffs(us, 2)
    host = OSQPLike().setup(P, q, A, l, u, eps_abs=1e-4, eps_rel=1e-4, warm_start=True)
    dq = DeviceQP(model.path, eps_abs=1e-4, eps_rel=1e-4).setup(P, q, model.path.assemble(us, 2))
    for it in (0, 2, 3):
        A, l, u = model.get_constraints_coeffs(us, it)
        host.update(l=l, u=u); host.update(Ax=A.data)
        dq.update(model.path.assemble(us, it))
        r0, r1 = host.solve(), dq.solve()
        assert r0.info.status == r1.info.status == 'solved'
        assert abs(r0.info.iter - r1.info.iter) <= 10
        assert np.max(np.abs(r0.x - r1.x)) < (1e-7 if it == 0 else 1e-4), it
        us = model.convert_us_vec_to_us_mat(r0.x[:36])


def test_car_scp_with_the_device_solver():
    """Car Model.define_problem(us, it, solver='device'): iteration 0 (relaxed pattern, a handful of rows) goes to
    the host solver, iterations >= 1 are solved on the device; same trajectory as the host path."""
    from riskaversetrajopt_b200.car.driving import Model
    opts = dict(eps_abs=1e-6, eps_rel=1e-6, polish=False, max_iter=200000)
    np.random.seed(0); host = Model(30, 'saa', 0.1)
    np.random.seed(0); dev = Model(30, 'saa', 0.1)
    us_h = us_d = host.initial_guess_us_mat()
    for it in range(5):
        host.define_problem(us_h, it, tail=False, solver_opts=opts); us_h, t_h = host.solve()
        dev.define_problem(us_d, it, solver='device', tail=False, solver_opts=opts); us_d, t_d = dev.solve()
        assert (dev._dqp is None) == (it == 0)
        assert dev.res.info.status == 'solved'
        assert np.max(np.abs(us_h - us_d)) < 1e-4 and abs(t_h - t_d) < 1e-4, it      # rounding differences grow along the SCP


def test_reset_forgets_the_warm_start():
    """DeviceQP.reset(): the next solve starts from zero like a freshly set-up solver (same iterates)."""
    from riskaversetrajopt_b200.device_qp import DeviceQP
    model = _drone_model(40)
    P, q = model.get_objective_coeffs()
    us = model.initial_guess_us_mat()
    b = model.path.assemble(us, 0)             # the relaxed first iteration: converges in tens of iterations
    a = DeviceQP(model.path, eps_abs=1e-4, eps_rel=1e-4).setup(P, q, b)
    r0 = a.solve()
    a.reset(rho=0.1)
    a.update(b)
    r1 = a.solve()
    assert r0.info.status == r1.info.status == 'solved' and r0.info.iter == r1.info.iter
    assert np.max(np.abs(r0.x - r1.x)) < 1e-9
