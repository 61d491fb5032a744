"""Quadrotor CVaR obstacle avoidance: host ``Model`` backed by the CUDA path.

Same constructor and method signatures as ``Model`` in the reference's
``drone/drone_risk.py:70-469`` (and its copy in ``drone_times.py``), so an SCP
driver written against the reference works unchanged.  The per-sample work –
rollout, control Jacobians, obstacle rows, sample means – runs in
``libsaa_b200.so``; nothing here computes on the CPU beyond O(n_u*S) glue.

Differences a caller can observe (all documented in INTEGRATION.md):
  * matrices are float64 NumPy/SciPy objects built from a *static, structural*
    CSC pattern (explicit zeros are kept; the reference's pattern comes from
    scanning a dense matrix and silently loses coincidental zeros);
  * ``get_all_constraints_coeffs_all`` densifies on the host and is meant for
    small M only – the dense matrix is O(M^2), which is why the reference cannot
    scale; use ``get_constraints_coeffs`` (sparse) instead;
  * no JAX arrays: inputs/outputs are NumPy.
"""
import numpy as np
import scipy.sparse as sp

from . import drone_params
from .. import _lib
from ..device_path import DevicePath

n_x, n_u = drone_params.n_x, drone_params.n_u
T, R = drone_params.T, drone_params.R
OSQP_TOL, OSQP_POLISH = drone_params.OSQP_TOL, drone_params.OSQP_POLISH


class Model:
    def __init__(self, S, DWs, masses, obs_Qs, method='saa', alpha=0.1, *,
                 variant='risk', precision='fp64', device=None, verbose=False, shard=None):
        """``shard=(M_global, sample_offset)``: these samples are one rank's block of a sharded set
        (``riskaversetrajopt_b200.dist``); the matrices then describe the rank's own row block."""
        if verbose:
            print("Initializing Model with")
            print("> method =", method)
            print("> alpha  =", alpha)
            print("> S      =", S)
        self.method, self.S, self.dt = method, int(S), T / S
        self.u_max, self.u_min = drone_params.u_max, -drone_params.u_max
        self.alpha = alpha
        self.beta = drone_params.beta
        self.drag_coefficient = drone_params.drag_coefficient
        self.DWs, self.masses, self.obs_Qs = DWs, masses, obs_Qs
        self.M = int(np.shape(masses)[0])
        self.variant = variant
        M_global, sample_offset = (self.M, 0) if shard is None else shard
        self.path = DevicePath(_lib.SAA_DRONE, method, self.S, alpha, self.M, M_global=M_global,
                               sample_offset=sample_offset, variant=variant, precision=precision, device=device)
        if shard is not None:
            self.path.set_output_geometry(self.M, 0)
        self.path.set_params_drone(drone_params, OSQP_TOL)
        self.path.set_samples_drone(masses, DWs, obs_Qs)
        self.osqp_prob = None

    # -- conversions (reference drone_risk.py:95-106) -----------------------------
    def convert_us_vec_to_us_mat(self, us_vec):
        return np.array(np.reshape(us_vec, (n_u, self.S), 'F').T)

    def convert_us_mat_to_us_jaxvec(self, us_mat):
        return np.reshape(us_mat, (self.S * n_u), 'C')

    # -- reference drone_risk.py:108-120; drone_times.py:144 fills all three axes ----
    def initial_guess_us_mat(self):
        us = np.zeros((self.S, n_u))
        guess = (self.u_max + self.u_min) / 2.0 + 1e-2
        if self.variant == 'times':
            us[:, :] = guess
        else:
            us[:, :(n_u - 1)] = guess
        return us

    # -- reference drone_risk.py:157-162 ------------------------------------------
    def us_to_state_trajectories(self, us_mat):
        return self.path.rollout(us_mat).cpu().numpy().astype(np.float64)

    # -- reference drone_risk.py:221-237 ------------------------------------------
    def get_control_constraints_coeffs_all(self):
        nu = n_u * self.S
        A = np.zeros((nu, nu + self.M + 2))
        A[np.arange(nu), np.arange(nu)] = 1.0
        return A, self.u_min * np.ones(nu), self.u_max * np.ones(nu)

    # -- reference drone_risk.py:376-399.  P, q are static; built sparse directly ----
    def get_objective_coeffs(self):
        nu, n = n_u * self.S, n_u * self.S + self.M + 2
        Rd = 2 * self.dt * np.asarray(R)
        rows, cols = np.nonzero(np.kron(np.eye(self.S), Rd))
        vals = np.kron(np.eye(self.S), Rd)[rows, cols]
        P = sp.csc_matrix((np.append(vals, 10000.0), (np.append(rows, n - 2), np.append(cols, n - 2))),
                          shape=(n, n))
        q = np.zeros(n)
        q[-2] = 10000.0
        return P, q

    # -- reference drone_risk.py:401-423: THE DROP-IN BOUNDARY --------------------------
    def get_constraints_coeffs(self, us_mat, scp_iter, copy=True):
        """-> (A: csc_matrix (68+61M, 62+M) for 'saa', l, u) with l <= A z <= u."""
        return self.path.csc(us_mat, scp_iter, copy=copy)

    # -- reference drone_risk.py:282-374 (dense; small M only) -----------------------
    def get_all_constraints_coeffs_all(self, us_mat):
        A, l, u = self.path.csc(us_mat, 2)
        nrow = A.shape[0] - n_u * self.S
        return A[:nrow].toarray(), l[:nrow], u[:nrow]

    # -- SCP glue (reference drone_risk.py:425-469), host solver = OSQP stand-in -----
    def define_problem(self, us_mat_p, verbose=False, solver=None, tail=None, solver_opts=None):
        """``tail``: None = automatic (the tail-reduced subproblem when M > 20 000, where no host QP
        ingests the full matrix), False = always the full problem as the reference, True / a margin
        (float) / dict(margin=, max_resolves=) = the tail-reduced subproblem (``tail_scp``); ``self.left_out_margin``
        reports after each solve whether the reduction was exact (<= 0)."""
        from ..qp import make_solver
        from .. import tail_scp
        scp_iter = 2
        self._tail = self._dqp = None
        if tail is None:
            tail = self.method == 'saa' and self.M > tail_scp.DEFAULT_TAIL_THRESHOLD
        if solver == 'device' and (tail is False or self.method != 'saa'):
            # solver='device': the QP is solved where the matrix is (device_qp.DeviceQP), nothing sample-sized
            # crosses PCIe; with tail=... it is the tail-reduced subproblem that is solved there (TailSCP)
            from ..device_qp import DeviceQP
            self.P, self.q = self.get_objective_coeffs()
            self._dqp = DeviceQP(self.path, **{**dict(eps_abs=OSQP_TOL, eps_rel=OSQP_TOL, polish=OSQP_POLISH, verbose=verbose), **(solver_opts or {})})
            self._dqp.setup(self.P, self.q, self.path.assemble(us_mat_p, scp_iter))
            self.osqp_prob = self._dqp
            return True
        if tail is not False and tail is not None and self.method == 'saa':
            opts = dict(tail) if isinstance(tail, dict) else {}
            margin = opts.get('margin', 0.25) if (tail is True or isinstance(tail, dict)) else float(tail)
            self._tail = tail_scp.TailSCP(self, n_u * self.S, OSQP_TOL, OSQP_POLISH, margin, solver, verbose,
                                               max_resolves=opts.get('max_resolves', 0), solver_opts=solver_opts)
            self._tail.define(us_mat_p, scp_iter)
            self.osqp_prob = self._tail.prob
            return True
        self.P, self.q = self.get_objective_coeffs()
        self.A, self.l, self.u = self.get_constraints_coeffs(us_mat_p, scp_iter)
        self.osqp_prob = make_solver(solver)
        self.osqp_prob.setup(self.P, self.q, self.A, self.l, self.u,
                             **{**dict(eps_abs=OSQP_TOL, eps_rel=OSQP_TOL, warm_start=True, verbose=verbose,
                                       polish=OSQP_POLISH), **(solver_opts or {})})
        return True

    def update_problem(self, us_mat_p, scp_iter=0, verbose=False):
        if getattr(self, '_dqp', None) is not None:
            self._dqp.update(self.path.assemble(us_mat_p, scp_iter))
            return True
        if getattr(self, '_tail', None) is not None:
            self._tail.update(us_mat_p, scp_iter)
            return True
        # the reference also rebuilds P, q here (:445) although they never change
        self.A, self.l, self.u = self.get_constraints_coeffs(us_mat_p, scp_iter)
        self.osqp_prob.update(l=self.l, u=self.u)
        self.osqp_prob.update(Ax=self.A.data)
        return True

    def solve(self, verbose=True):
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
            return self.convert_us_vec_to_us_mat(self.res.x[:(n_u * self.S)]), self.res.x[-1]
        self.res = self.osqp_prob.solve()
        if self.res.info.status != 'solved':
            print("[solve]: Problem infeasible.")
        us_sol = self.convert_us_vec_to_us_mat(self.res.x[:(n_u * self.S)])
        ys, t_risk_sol = self.res.x[(n_u * self.S):-2], self.res.x[-1]
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

    # -- Monte-Carlo verification (reference drone_risk.py:656-662, :694) -------------
    def monte_carlo_constraints(self, us_mat):
        """-> (B_satisfied (M,) bool, max_constraint (M,)) as the vmapped
        ``monte_carlo_no_collisions_constraint_verification``."""
        Z, _ = self.path.cvar_terms(us_mat, 0.0, 1e-6)
        Z = Z.cpu().numpy().astype(np.float64)
        return Z <= 1e-6, Z

    def monte_carlo_avar(self, us_mat, t_risk, alpha=None):
        """t + mean(max(Z - t, 0)) / alpha (closed form the reference evaluates at :694)."""
        alpha = self.alpha if alpha is None else alpha
        _, out3 = self.path.cvar_terms(us_mat, t_risk, 1e-6, want_Z=False)
        return t_risk + float(out3[0].item()) / (self.M * alpha)


def L2_error_us(us_mat, us_mat_prev):
    """reference drone_risk.py:471-476"""
    error = np.mean(np.linalg.norm(us_mat - us_mat_prev, axis=-1))
    return error / np.mean(np.linalg.norm(us_mat, axis=-1))
