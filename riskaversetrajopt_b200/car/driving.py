"""Ego car vs. uncertain pedestrian (CVaR separation constraint): host ``Model``
backed by the CUDA path.

Same constructor and method signatures as ``Model`` in the reference's
``car/driving.py:83-456``: ``Model(M, method='saa', alpha=0.05)`` draws the
uncertain parameters from the legacy global NumPy stream in the reference's
order, ``define_problem(us, scp_iter)`` / ``solve()`` drive the SCP loop, and
``get_constraints_coeffs(us_mat, scp_iter) -> (A csc, l, u)`` is the drop-in
boundary.  The per-sample work runs in ``libsaa_b200.so``.

Reference quirks that are reproduced on purpose:
  * the pedestrian force adds the scalar ``w_s (1.3 - x[7])`` to both components
    (car/driving.py:150-157);
  * Brownian increments are scaled by sqrt(dt) twice (:116, :200);
  * at ``scp_iter == 0`` rows ``>= n_x = 8`` are multiplied by exactly 0 although
    there are only 4 final rows, so the CVaR row and the first three ``-y_i`` rows
    survive, every other risk row vanishes from the CSC pattern, and the lower
    bounds of the zeroed rows are ``-inf * 0 = nan`` (:411-415).
"""
import numpy as np
import scipy.sparse as sp

from . import driving_params as p
from .. import _lib
from ..device_path import DevicePath

n_x, n_u, S, dt = p.n_x, p.n_u, p.S, p.dt
OSQP_TOL, OSQP_POLISH = p.OSQP_TOL, p.OSQP_POLISH
BETA = 3e-2     # diffusion magnitude, reference car/driving.py:94


def sample_uncertain_parameters(M, method='saa'):
    """Consumes the global legacy RNG exactly like the reference constructor
    (car/driving.py:95-120): M uniforms (omega_speed), M uniforms
    (omega_repulsive), [saa] M x randn(4) for the pedestrian's initial state,
    then M*S*n_x normals.  -> (states_init, omegas_speed, omegas_repulsive, DWs)."""
    w_s = np.random.uniform(p.omega_speed_nom - p.omega_speed_del,
                            p.omega_speed_nom + p.omega_speed_del, M)
    w_r = np.random.uniform(p.omega_repulsive_nom - p.omega_repulsive_del,
                            p.omega_repulsive_nom + p.omega_repulsive_del, M)
    states_init = np.repeat(np.asarray(p.state_init, dtype=np.float64)[None, :], M, axis=0)
    if method == 'saa':
        std = np.sqrt(np.diag(p.variance_ped_initial_state))
        states_init[:, 4:] += np.random.randn(M, 4) * std[None, :]
    DWs = np.sqrt(dt) * np.random.randn(M, S, n_x)
    if method == 'baseline':
        DWs, w_s, w_r = 0 * DWs, 0 * w_s, 0 * w_r
    return states_init, w_s, w_r, DWs


class Model:
    def __init__(self, M, method='saa', alpha=0.05, *, samples=None, precision='fp64',
                 device=None, verbose=False):
        if verbose:
            print("Initializing Model with")
            print("> method =", method)
            print("> alpha  =", alpha)
        self.method, self.alpha, self.M = method, alpha, int(M)
        self.u_max, self.u_min = p.u_max, -p.u_max
        self.beta = BETA
        if samples is None:
            samples = sample_uncertain_parameters(self.M, method)
        self.states_init, self.omegas_speed, self.omegas_repulsive, self.DWs = samples
        self.path = DevicePath(_lib.SAA_CAR, method, S, alpha, self.M, precision=precision,
                               device=device)
        self.path.set_params_car(p, self.beta, OSQP_TOL)
        self.path.set_samples_car(self.states_init, self.omegas_speed, self.omegas_repulsive,
                                  self.DWs)
        self.osqp_prob = None

    # -- conversions (reference car/driving.py:122-130) ------------------------------
    def convert_us_vec_to_us_mat(self, us_vec):
        return np.array(np.reshape(us_vec, (n_u, S), 'F').T)

    def convert_us_mat_to_us_jaxvec(self, us_mat):
        return np.reshape(us_mat, (S * n_u), 'C')

    def initial_guess_us_mat(self):                       # :132-143
        return np.full((S, n_u), (self.u_max + self.u_min) / 2.0 + 1e-2)

    def us_to_state_trajectories(self, us_mat):           # :207-214
        return self.path.rollout(us_mat).cpu().numpy().astype(np.float64)

    def get_control_constraints_coeffs_all(self):         # :243-258
        nu = n_u * S
        A = np.zeros((nu, nu + self.M + 2))
        A[np.arange(nu), np.arange(nu)] = 1.0
        return A, self.u_min * np.ones(nu), self.u_max * np.ones(nu)

    def get_objective_coeffs(self):                       # :375-397 (static, built sparse)
        n = n_u * S + self.M + 2
        Rd = 2 * dt * np.asarray(p.R)
        D = np.kron(np.eye(S), Rd)
        rows, cols = np.nonzero(D)
        P = sp.csc_matrix((np.append(D[rows, cols], 1000.0),
                           (np.append(rows, n - 2), np.append(cols, n - 2))), shape=(n, n))
        q = np.zeros(n)
        q[-2] = 1000.0
        return P, q

    # -- reference car/driving.py:399-421: THE DROP-IN BOUNDARY -------------------------
    def get_constraints_coeffs(self, us_mat, scp_iter, copy=True):
        """-> (A csc (46+21M, 42+M), l, u).  At scp_iter 0 the matrix has the reduced
        pattern described in the module docstring."""
        return self.path.csc(us_mat, scp_iter, copy=copy)

    def get_all_constraints_coeffs_all(self, us_mat):     # :302-373 (dense; small M only)
        A, l, u = self.path.csc(us_mat, 1)
        nrow = A.shape[0] - n_u * S
        return A[:nrow].toarray(), l[:nrow], u[:nrow]

    # -- SCP glue (reference :423-456) ---------------------------------------------------
    def define_problem(self, us_mat_p, scp_iter=0, verbose=False, solver=None, tail=None, solver_opts=None):
        """``tail`` as in the drone ``Model.define_problem``: None = automatic (tail-reduced subproblem from
        scp_iter >= 1 on when M > 20 000), False = the full problem as the reference, True / margin = tail."""
        from ..qp import make_solver
        from .. import tail_scp
        if tail is None:
            tail = self.method == 'saa' and self.M > tail_scp.DEFAULT_TAIL_THRESHOLD
        if solver == 'device' and scp_iter == 0:
            solver = None                   # the relaxed problem of scp_iter 0 (a handful of rows, driving.py:411-415)
        if solver == 'device' and tail is False:
            # the QP is solved where the matrix is (device_qp.DeviceQP); with tail=... the tail-reduced
            # subproblem is solved there instead (TailSCP)
            if self.method == 'saa':
                from ..device_qp import DeviceQP
                b = self.path.assemble(us_mat_p, scp_iter)
                if getattr(self, '_dqp', None) is None:
                    self.P, self.q = self.get_objective_coeffs()
                    self._dqp = DeviceQP(self.path, **{**dict(eps_abs=OSQP_TOL, eps_rel=OSQP_TOL, polish=OSQP_POLISH, verbose=verbose), **(solver_opts or {})})
                    self._dqp.setup(self.P, self.q, b)
                else:
                    self._dqp.update(b)
                self._tail, self.osqp_prob = None, self._dqp
                return True
            solver = None
        self._dqp = None
        if tail is not False and tail is not None and self.method == 'saa' and scp_iter >= 1:
            if scp_iter == 1 or getattr(self, '_tail', None) is None:
                opts = dict(tail) if isinstance(tail, dict) else {}
                margin = opts.get('margin', 0.25) if (tail is True or isinstance(tail, dict)) else float(tail)
                self._tail = tail_scp.TailSCP(self, n_u * S, OSQP_TOL, OSQP_POLISH, margin, solver, verbose,
                                               max_resolves=opts.get('max_resolves', 0), solver_opts=solver_opts)
                self._tail.define(us_mat_p, scp_iter)
            else:
                self._tail.update(us_mat_p, scp_iter)
            self.osqp_prob = self._tail.prob
            return True
        self._tail = None
        self.P, self.q = self.get_objective_coeffs()
        self.A, self.l, self.u = self.get_constraints_coeffs(us_mat_p, scp_iter)
        if scp_iter == 0 or scp_iter == 1:
            # the pattern changes between iteration 0 and 1: set up again (:429-437)
            self.osqp_prob = make_solver(solver)
            l = np.where(np.isnan(self.l), -np.inf, self.l)   # OSQP rejects nan; rows are empty anyway
            self.osqp_prob.setup(self.P, self.q, self.A, l, self.u,
                                 **{**dict(eps_abs=OSQP_TOL, eps_rel=OSQP_TOL, warm_start=True, verbose=verbose,
                                           polish=OSQP_POLISH), **(solver_opts or {})})
        else:
            self.osqp_prob.update(l=self.l, u=self.u)
            self.osqp_prob.update(Ax=self.A.data)
        return True

    def solve(self, verbose=False):
        if getattr(self, '_dqp', None) is not None:
            self.res = self._dqp.solve()
            if self.res.info.status != 'solved':
                print("[solve]: Problem infeasible.")
            return self.convert_us_vec_to_us_mat(self.res.u), self.res.t
        if getattr(self, '_tail', None) is not None:
            self.res, self.left_out_margin = self._tail.solve()
            self.osqp_prob = self._tail.prob
            if self.res.info.status != 'solved':
                print("[solve]: Problem infeasible.")
            return self.convert_us_vec_to_us_mat(self.res.x[:(n_u * S)]), self.res.x[-1]
        self.res = self.osqp_prob.solve()
        if self.res.info.status != 'solved':
            print("[solve]: Problem infeasible.")
        us_sol = self.convert_us_vec_to_us_mat(self.res.x[:(n_u * S)])
        ys, t_risk_sol = self.res.x[(n_u * S):-2], self.res.x[-1]
        if verbose:
            print("y_min =", np.min(ys))
            print("slack_var =", self.res.x[-2])
        return us_sol, t_risk_sol

    # -- tail-reduced subproblem (no reference counterpart; riskaversetrajopt_b200/tail.py) ----
    def tail_subproblem(self, K=None, margin=0.25):
        """-> ``TailSubproblem``: the QP restricted to the K samples with the largest constraint
        values at the iterate (default K = ceil((1 + margin) alpha M)), selected on the device."""
        from ..tail import TailSubproblem
        return TailSubproblem(self.path, K=K, margin=margin)

    # -- Monte-Carlo verification (reference :630-638, :670) -------------------------------
    def monte_carlo_constraints(self, us_mat):
        Z, _ = self.path.cvar_terms(us_mat, 0.0, 1e-6)
        Z = Z.cpu().numpy().astype(np.float64)
        return Z <= 1e-6, Z

    def monte_carlo_avar(self, us_mat, t_risk, alpha=None):
        alpha = self.alpha if alpha is None else alpha
        _, out3 = self.path.cvar_terms(us_mat, t_risk, 1e-6, want_Z=False)
        return t_risk + float(out3[0].item()) / (self.M * alpha)


def L2_error_us(us_mat, us_mat_prev):
    """reference car/driving.py:458-464"""
    error = np.mean(np.linalg.norm(us_mat - us_mat_prev, axis=-1))
    return error / np.mean(np.linalg.norm(us_mat, axis=-1))
