"""SCP glue for the tail-reduced subproblem: what ``Model.define_problem / update_problem / solve`` do
when the sample set is too large for a host QP (``tail=...``).

The reference hands the full ``(68 + 61 M) x (62 + M)`` matrix to OSQP every iteration
(drone/drone_risk.py:425-452).  At M >= ~10^4 no host solver ingests that, and bringing it to the
host is PCIe bound.  Here the QP is restricted to the K samples with the largest constraint values
at the iterate (``tail.TailSubproblem``: selected on the device, same pattern every iteration).
After each solve the samples that were left out are checked against the new risk level
(``left_out_margin(t) = max_{i left out} Z_i - t`` with Z_i at the linearisation point); the value is
reported as ``Model.left_out_margin``.  A non-positive margin certifies that the reduced QP had the
full QP's minimiser.  The check is conservative far from convergence -- at the reference's initial
guess EVERY sample violates the obstacle constraints, so the "tail" is the whole set and the margin is
positive no matter how large K is -- therefore growing K is opt-in: with ``max_resolves > 0`` a
positive margin doubles K, re-defines the subproblem at the same iterate and re-solves, at most that
many times per SCP iteration.  Near convergence (where <= alpha M samples are active) the margin turns
non-positive and the SCP continues on the exact reduction.
"""
import numpy as np

DEFAULT_TAIL_THRESHOLD = 20000      # Model(...).define_problem: M above this switches to the tail subproblem


class TailSCP:
    def __init__(self, model, n_ctrl, osqp_tol, polish, margin=0.25, solver=None, verbose=False, max_resolves=0,
                 solver_opts=None):
        self.model, self.nu = model, int(n_ctrl)
        self.osqp_tol, self.polish, self.solver_name, self.verbose = osqp_tol, polish, solver, verbose
        self.margin, self.max_resolves = margin, int(max_resolves)
        self.solver_opts = dict(solver_opts or {})       # overrides of eps_abs / eps_rel / polish / max_iter ...
        self.tail = None
        self.resolves = 0

    def _define(self, us_mat, scp_iter, K=None):
        from .qp import make_solver
        m = self.model
        self.tail = m.tail_subproblem(K=K, margin=self.margin)
        self.P, self.q = self.tail.get_objective_coeffs(*m.get_objective_coeffs())
        if self.solver_name == 'device':
            # the reduced QP is solved where it was assembled: only (u, slack, t) come back
            from .device_qp import DeviceQP
            kw = dict(eps_abs=self.osqp_tol, eps_rel=self.osqp_tol, polish=self.polish, verbose=self.verbose)
            kw.update(self.solver_opts)
            self.prob = _DeviceProb(DeviceQP(self.tail.sub, **kw), self.tail)
            self.prob.setup(self.P, self.q, us_mat, scp_iter)
            self._last = (np.array(us_mat, dtype=np.float64), scp_iter)
            return
        self.A, self.l, self.u, self.idx = self.tail.get_constraints_coeffs(us_mat, scp_iter)
        self.prob = make_solver(self.solver_name)
        kw = dict(eps_abs=self.osqp_tol, eps_rel=self.osqp_tol, warm_start=True, verbose=self.verbose, polish=self.polish)
        kw.update(self.solver_opts)
        self.prob.setup(self.P, self.q, self.A, self.l, self.u, **kw)
        self._last = (np.array(us_mat, dtype=np.float64), scp_iter)

    def define(self, us_mat, scp_iter):
        self._define(us_mat, scp_iter)

    def update(self, us_mat, scp_iter):
        if self.solver_name == 'device':
            self.prob.update(us_mat, scp_iter)
            self._last = (np.array(us_mat, dtype=np.float64), scp_iter)
            return
        self.A, self.l, self.u, self.idx = self.tail.get_constraints_coeffs(us_mat, scp_iter, copy=False)
        self.prob.update(l=self.l, u=self.u)
        self.prob.update(Ax=self.A.data)
        self._last = (np.array(us_mat, dtype=np.float64), scp_iter)

    def solve(self):
        """-> (res, margin); with ``max_resolves > 0`` re-solves with a doubled K while a left-out sample
        would be active."""
        tries = 0
        while True:
            res = self.prob.solve()
            t_risk = float(res.x[-1])
            if self._last[1] < self.model.path.relax_threshold():
                # relaxed first iterations (drone_risk.py:413-417, driving.py:411-415): the risk rows are
                # scaled away, t is not a risk level yet -- nothing to check
                return res, -np.inf
            margin = self.tail.left_out_margin(t_risk)
            if margin <= 0 or self.tail.K >= self.model.M or tries >= self.max_resolves:
                return res, margin
            tries += 1
            self.resolves += 1
            self._define(self._last[0], self._last[1], K=min(self.model.M, 2 * self.tail.K))


class _DeviceProb:
    """``DeviceQP`` on the tail-reduced subproblem, with the solve() result shaped like OSQP's (x = (u, ..., t))."""

    def __init__(self, dqp, tail):
        self.dqp, self.tail = dqp, tail

    def setup(self, P, q, us_mat, scp_iter):
        self.dqp.setup(P, q, self.tail.assemble(us_mat, scp_iter))

    def update(self, us_mat, scp_iter):
        # the selection changes with the iterate: sample k of the subproblem is another sample now.  The
        # multipliers / y of the previous iterate are kept as a warm start all the same (most of the tail persists).
        self.dqp.update(self.tail.assemble(us_mat, scp_iter))

    def solve(self):
        return self.dqp.solve()
