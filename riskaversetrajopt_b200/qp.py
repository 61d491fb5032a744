"""Host QP solvers with the slice of the OSQP Python interface the reference uses
(``setup`` / ``update(l=, u=)`` / ``update(Ax=)`` / ``solve`` -> ``res.x``,
``res.info.status``; drone/drone_risk.py:433-461, car/driving.py:429-447).

The convex solve is NOT part of the accelerated path (BASELINE.json: "the convex solve
stays on the reference's host solver and is timed separately as context").  OSQP is not
installed in this image, so two stand-ins are provided:

* ``'admm'``  – a from-scratch implementation of the OSQP algorithm (Stellato et al.,
  2020: ADMM, Ruiz equilibration, per-constraint and adaptive rho, warm start, solution
  polishing on the guessed active set with iterative refinement); because the
  SAA programs have m ~ 61 n the KKT system is solved in its reduced n x n form
  (dense Cholesky / sparse LU), 2-3x faster than factorising the (n + m) system.
* ``'highs'`` – HiGHS' QP active-set solver through SciPy's bundled (private) binding;
  exact, used to cross-check the ADMM solution on small problems.

If the real ``osqp`` package is importable, ``make_solver('osqp')`` returns it.
"""
import time
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_INF = 1e20


def _bounds(l, u):
    """Bounds as the solver uses them.  NaN bounds (the reference's car problem at scp_iter 0 has
    ``-inf * 0 = nan`` lower bounds on rows whose entries were zeroed, car/driving.py:411-415; OSQP
    rejects NaN) are treated as "no bound": the rows are empty, so any finite value is equivalent."""
    l = np.asarray(l, dtype=np.float64) if l is not None else None
    u = np.asarray(u, dtype=np.float64) if u is not None else None
    if l is not None:
        l = np.maximum(np.where(np.isnan(l), -_INF, l), -_INF)
    if u is not None:
        u = np.minimum(np.where(np.isnan(u), _INF, u), _INF)
    return l, u


def make_solver(name=None):
    name = name or 'admm'
    if name == 'osqp':
        import osqp
        return osqp.OSQP()
    if name == 'highs':
        return HighsQP()
    if name == 'admm':
        return OSQPLike()
    raise ValueError(f"unknown QP solver {name!r}")


class OSQPLike:
    """ADMM QP solver following the OSQP algorithm; same call pattern as ``osqp.OSQP``."""

    def setup(self, P, q, A, l, u, eps_abs=1e-3, eps_rel=1e-3, max_iter=20000, rho=0.1, sigma=1e-6,
              alpha=1.6, warm_start=True, verbose=False, polish=False, scaling=10,
              adaptive_rho_interval=50, check_interval=10, **_ignored):
        self.n, self.m = P.shape[0], A.shape[0]
        self.P = sp.csc_matrix(P, dtype=np.float64)
        self.P = sp.triu(self.P, format='csc') + sp.triu(self.P, 1, format='csc').T   # symmetric
        self.A = sp.csc_matrix(A, dtype=np.float64).copy()
        self.q = np.asarray(q, dtype=np.float64).copy()
        self.l, self.u = _bounds(l, u)
        # polish=True (the reference's setting, drone_risk.py:436-440) is EMULATED by iterating to 1e-6 with a
        # larger iteration cap: the SCP of the reference does not converge when its QPs are only solved to
        # 3e-4 (examples/car_scp.py).  polish='kkt' runs OSQP's actual polishing step after ADMM reaches
        # (eps_abs, eps_rel) -- the KKT system of the guessed active set with iterative refinement
        # (``_polish``) -- and falls back to the emulation when the polished point does not improve both
        # residuals.  On the CVaR programs that happens on every unrelaxed SCP iteration (the active set
        # guessed at eps = 1e-3 is wrong: the program is degenerate), and mixing polished and unpolished
        # iterations perturbs the trust-region-free SCP, so it is opt-in.  Infeasible / unbounded programs end
        # with OSQP's status strings 'primal infeasible' / 'dual infeasible' (``_certificates``).
        self.polish = polish == 'kkt'
        if polish and not self.polish:
            eps_abs, eps_rel, max_iter = min(eps_abs, 1e-6), min(eps_rel, 1e-6), max(max_iter, 200000)
        self.opts = SimpleNamespace(eps_abs=eps_abs, eps_rel=eps_rel, max_iter=max_iter, rho=rho,
                                    sigma=sigma, alpha=alpha, warm_start=warm_start, verbose=verbose,
                                    scaling=scaling, adaptive_rho_interval=adaptive_rho_interval,
                                    check_interval=check_interval)
        self.x = np.zeros(self.n)
        self.z = np.zeros(self.m)
        self.y = np.zeros(self.m)
        self.rho = rho
        self._scale()
        self._factor()
        return self

    # -- Ruiz equilibration of [[P, A^T], [A, 0]] ------------------------------------------
    def _scale(self):
        n, m = self.n, self.m
        D, E, c = np.ones(n), np.ones(m), 1.0
        P, A, q = self.P.copy(), self.A.copy(), self.q.copy()
        for _ in range(self.opts.scaling):
            col_P = _col_inf_norm(P)
            col_A = _col_inf_norm(A)
            row_A = _row_inf_norm(A)
            d = 1.0 / np.sqrt(_guard(np.maximum(col_P, col_A)))
            e = 1.0 / np.sqrt(_guard(row_A))
            P = sp.diags(d) @ P @ sp.diags(d)
            A = sp.diags(e) @ A @ sp.diags(d)
            q = d * q
            D, E = D * d, E * e
            g = 1.0 / max(np.mean(_col_inf_norm(P)), np.max(np.abs(q)), 1e-4)
            g = min(max(g, 1e-4), 1e4)
            P, q, c = P * g, q * g, c * g
        self.D, self.E, self.c = D, E, c
        self.Ps, self.As, self.qs = sp.csc_matrix(P), sp.csc_matrix(A), q
        self.ls, self.us = self.E * self.l, self.E * self.u
        self.ls[self.l <= -_INF], self.us[self.u >= _INF] = -_INF, _INF

    def _rho_vec(self):
        eq = np.abs(self.us - self.ls) < 1e-10
        free = (self.ls <= -_INF) & (self.us >= _INF)
        r = np.full(self.m, self.rho)
        r[eq] = 1e3 * self.rho
        r[free] = 1e-6
        return r

    def _factor(self):
        """The SAA programs have far more constraints than variables (m ~ 61 n), so the ADMM step
        solves the reduced system (P + sigma I + A^T R A) x = sigma x_k - q + A^T (R z_k - y_k),
        nu = R (A x - z_k) + y_k, with R = diag(rho) -- the quasi-definite (n + m) KKT system of
        the OSQP paper with nu eliminated.  Dense Cholesky while n is small, sparse LU otherwise."""
        self.rho_v = self._rho_vec()
        S = (self.Ps + self.opts.sigma * sp.identity(self.n)
             + self.As.T @ sp.diags(self.rho_v) @ self.As)
        if self.n <= 6000:
            import scipy.linalg as sla
            c = sla.cho_factor(S.toarray(), lower=True, check_finite=False)
            self._solve_S = lambda rhs: sla.cho_solve(c, rhs, check_finite=False)
        else:
            lu = spla.splu(sp.csc_matrix(S))
            self._solve_S = lu.solve

    # -- OSQP-style updates -------------------------------------------------------------------
    def update(self, q=None, l=None, u=None, Ax=None, Px=None, **_ignored):
        if q is not None:
            self.q = np.asarray(q, dtype=np.float64).copy()
            self.qs = self.c * self.D * self.q
        if l is not None:
            self.l = _bounds(l, None)[0]
        if u is not None:
            self.u = _bounds(None, u)[1]
        if l is not None or u is not None:
            eq_before = np.abs(self.us - self.ls) < 1e-10
            self.ls, self.us = self.E * self.l, self.E * self.u
            self.ls[self.l <= -_INF], self.us[self.u >= _INF] = -_INF, _INF
            if np.any(eq_before != (np.abs(self.us - self.ls) < 1e-10)):
                self._factor()
        if Ax is not None:
            Ax = np.asarray(Ax, dtype=np.float64)
            if Ax.shape != self.A.data.shape:
                raise ValueError("update(Ax=...) needs the values in the CSC order given to setup()")
            self.A.data[:] = Ax
            self.As = sp.csc_matrix(sp.diags(self.E) @ self.A @ sp.diags(self.D))
            self._factor()

    def warm_start(self, x=None, y=None):
        if x is not None:
            self.x = np.asarray(x, dtype=np.float64) / self.D
            self.z = self.As @ self.x
        if y is not None:
            self.y = self.c * np.asarray(y, dtype=np.float64) / self.E

    def solve(self):
        o = self.opts
        t0 = time.perf_counter()
        if not o.warm_start:
            self.x[:], self.z[:], self.y[:] = 0.0, 0.0, 0.0
        status, it = self._admm(o.eps_abs, o.eps_rel, o.max_iter)
        self.polished = False
        if self.polish and status == 'solved':
            self.polished = self._polish()
            if not self.polished and (o.eps_abs > 1e-6 or o.eps_rel > 1e-6):
                status, it2 = self._admm(min(o.eps_abs, 1e-6), min(o.eps_rel, 1e-6), max(o.max_iter, 200000))
                it += it2
        xs = self.D * self.x
        info = SimpleNamespace(status=status, iter=it, run_time=time.perf_counter() - t0, polished=self.polished,
                               obj_val=float(0.5 * xs @ (self.P @ xs) + self.q @ xs))
        return SimpleNamespace(x=xs, y=self.E * self.y / self.c, info=info)

    def _polish(self, delta=1e-6, refine=5):
        """OSQP's polishing: guess the active constraints from (z, y), solve
        [[P + delta I, A_act'], [A_act, -delta I]] (x, y_act) = (-q, b_act) with iterative refinement against the
        unregularised system, accept if both residuals improve.  Scaled quantities throughout."""
        x, z, y, ls, us = self.x, self.z, self.y, self.ls, self.us
        eq = np.abs(us - ls) < 1e-10
        low = ((z - ls) < -y) | eq
        upp = ((us - z) < y) & ~low
        idx = np.flatnonzero(low | upp)
        n, ma = self.n, idx.size
        b = np.where(low, ls, us)[idx]
        Aa = sp.csr_matrix(self.As)[idx]
        Kreg = sp.bmat([[self.Ps + delta * sp.identity(n), Aa.T], [Aa, -delta * sp.identity(ma)]], format='csc')
        Ktrue = sp.bmat([[self.Ps, Aa.T], [Aa, None]], format='csr') if ma else sp.csr_matrix(self.Ps)
        rhs = np.concatenate([-self.qs, b])
        try:
            lu = spla.splu(Kreg)
        except RuntimeError:
            return False
        sol = lu.solve(rhs)
        for _ in range(refine):
            sol = sol + lu.solve(rhs - Ktrue @ sol)
        xp, ya = sol[:n], sol[n:]
        if not np.all(np.isfinite(sol)):
            return False
        yp = np.zeros(self.m); yp[idx] = ya
        Ax = self.As @ xp
        zp = np.clip(Ax, ls, us)
        rp0, rd0, _, _ = self._residuals(x, z, y)
        rp1, rd1, _, _ = self._residuals(xp, zp, yp)
        if (rp1 < rp0 and rd1 < rd0) or (rp1 < 1e-10 and rd1 < 1e-10):
            self.x, self.z, self.y = xp, zp, yp
            return True
        return False

    def _certificates(self, dx, dy, eps_inf=1e-4):
        """OSQP's infeasibility certificates from the change of the iterates over one step (Stellato et al.
        2020, section 3.4), on unscaled quantities: 'primal infeasible' if dy is (almost) in the null space of
        A' and has negative support-function value on [l, u]; 'dual infeasible' if dx is a direction of
        unbounded descent that keeps A dx inside the recession cone of [l, u]."""
        dyu = self.E * dy / self.c
        ndy = np.max(np.abs(dyu), initial=0.0)
        if ndy > 1e-30:
            up, lo = np.maximum(dyu, 0.0), np.minimum(dyu, 0.0)
            if not (np.any((self.u >= _INF) & (up > eps_inf * ndy)) or np.any((self.l <= -_INF) & (lo < -eps_inf * ndy))):
                fu, fl = np.where(self.u >= _INF, 0.0, self.u), np.where(self.l <= -_INF, 0.0, self.l)
                if (np.max(np.abs(self.A.T @ dyu), initial=0.0) <= eps_inf * ndy
                        and fu @ up + fl @ lo <= -eps_inf * ndy):
                    return 'primal infeasible'
        dxu = self.D * dx
        ndx = np.max(np.abs(dxu), initial=0.0)
        if ndx > 1e-30:
            Adx = self.A @ dxu
            ok_rows = np.all(np.where(self.u >= _INF, True, Adx <= eps_inf * ndx)
                             & np.where(self.l <= -_INF, True, Adx >= -eps_inf * ndx))
            if (ok_rows and np.max(np.abs(self.P @ dxu), initial=0.0) <= eps_inf * ndx
                    and self.q @ dxu <= -eps_inf * ndx):
                return 'dual infeasible'
        return None

    def _admm(self, eps_abs, eps_rel, max_iter):
        o = self.opts
        x, z, y = self.x, self.z, self.y
        status, it = 'maximum iterations reached', 0
        for it in range(1, max_iter + 1):
            if it % o.check_interval == 0:
                x_prev, y_prev = x, y
            xt = self._solve_S(o.sigma * x - self.qs + self.As.T @ (self.rho_v * z - y))
            zt = self.As @ xt                      # = z + (nu - y) / rho with nu = rho (A xt - z) + y
            x = o.alpha * xt + (1 - o.alpha) * x
            zr = o.alpha * zt + (1 - o.alpha) * z
            z_new = np.clip(zr + y / self.rho_v, self.ls, self.us)
            y = y + self.rho_v * (zr - z_new)
            z = z_new
            if it % o.check_interval == 0 or it == max_iter:
                rp, rd, ep, ed = self._residuals(x, z, y, eps_abs, eps_rel)
                if rp <= ep and rd <= ed:
                    status = 'solved'
                    break
                if it % o.check_interval == 0 and it > 5 * o.check_interval:
                    cert = self._certificates(x - x_prev, y - y_prev)
                    if cert is not None:
                        status = cert
                        break
                if o.adaptive_rho_interval and it % o.adaptive_rho_interval == 0:
                    num = rp / max(ep, 1e-30)
                    den = rd / max(ed, 1e-30)
                    new_rho = float(np.clip(self.rho * np.sqrt(num / max(den, 1e-30)), 1e-6, 1e6))
                    if new_rho > 5 * self.rho or new_rho < self.rho / 5:
                        self.rho = new_rho
                        self._factor()
        self.x, self.z, self.y = x, z, y
        return status, it

    def _residuals(self, x, z, y, eps_abs=None, eps_rel=None):
        o = SimpleNamespace(eps_abs=self.opts.eps_abs if eps_abs is None else eps_abs,
                            eps_rel=self.opts.eps_rel if eps_rel is None else eps_rel)
        Einv, Dinv = 1.0 / self.E, 1.0 / self.D
        Ax = self.As @ x
        Px = self.Ps @ x
        Aty = self.As.T @ y
        rp = np.max(np.abs(Einv * (Ax - z))) if self.m else 0.0
        rd = np.max(np.abs(Dinv * (Px + self.qs + Aty))) / self.c
        ep = o.eps_abs + o.eps_rel * max(np.max(np.abs(Einv * Ax), initial=0.0), np.max(np.abs(Einv * z), initial=0.0))
        ed = o.eps_abs + o.eps_rel * max(np.max(np.abs(Dinv * Px)), np.max(np.abs(Dinv * Aty), initial=0.0),
                                         np.max(np.abs(Dinv * self.qs))) / self.c
        return rp, rd, ep, ed


def _guard(v):
    v = np.asarray(v, dtype=np.float64).copy()
    v[v < 1e-4] = 1.0
    return np.minimum(v, 1e4)


def _col_inf_norm(M):
    M = sp.csc_matrix(M)
    out = np.zeros(M.shape[1])
    if M.nnz:
        np.maximum.at(out, np.repeat(np.arange(M.shape[1]), np.diff(M.indptr)), np.abs(M.data))
    return out


def _row_inf_norm(M):
    M = sp.csr_matrix(M)
    out = np.zeros(M.shape[0])
    if M.nnz:
        np.maximum.at(out, np.repeat(np.arange(M.shape[0]), np.diff(M.indptr)), np.abs(M.data))
    return out


class HighsQP:
    """HiGHS QP (active set) through ``scipy.optimize._highspy`` -- private SciPy API, exact."""

    def setup(self, P, q, A, l, u, verbose=False, **_ignored):
        self.P = sp.csc_matrix(sp.tril(sp.csc_matrix(P, dtype=np.float64)))
        self.A = sp.csc_matrix(A, dtype=np.float64).copy()
        self.q = np.asarray(q, dtype=np.float64).copy()
        self.l, self.u = np.asarray(l, dtype=np.float64).copy(), np.asarray(u, dtype=np.float64).copy()
        self.verbose = verbose
        return self

    def update(self, q=None, l=None, u=None, Ax=None, **_ignored):
        if q is not None:
            self.q = np.asarray(q, dtype=np.float64).copy()
        if l is not None:
            self.l = np.asarray(l, dtype=np.float64).copy()
        if u is not None:
            self.u = np.asarray(u, dtype=np.float64).copy()
        if Ax is not None:
            self.A.data[:] = np.asarray(Ax, dtype=np.float64)

    def solve(self):
        from scipy.optimize._highspy import _core as hs
        t0 = time.perf_counter()
        n, m = self.A.shape[1], self.A.shape[0]
        inf = hs.kHighsInf
        h = hs._Highs()
        h.setOptionValue("output_flag", bool(self.verbose))
        lp = hs.HighsLp()
        lp.num_col_, lp.num_row_ = n, m
        lp.col_cost_ = self.q
        lp.col_lower_, lp.col_upper_ = np.full(n, -inf), np.full(n, inf)
        lp.row_lower_ = np.where(np.isfinite(self.l), self.l, -inf)
        lp.row_upper_ = np.where(np.isfinite(self.u), self.u, inf)
        lp.a_matrix_.format_ = hs.MatrixFormat.kColwise
        lp.a_matrix_.start_ = self.A.indptr.astype(np.int32)
        lp.a_matrix_.index_ = self.A.indices.astype(np.int32)
        lp.a_matrix_.value_ = self.A.data
        h.passModel(lp)
        hess = hs.HighsHessian()
        hess.dim_ = n
        hess.format_ = hs.HessianFormat.kTriangular
        hess.start_ = self.P.indptr.astype(np.int32)
        hess.index_ = self.P.indices.astype(np.int32)
        hess.value_ = self.P.data
        h.passHessian(hess)
        h.run()
        ok = h.getModelStatus() == hs.HighsModelStatus.kOptimal
        x = np.array(h.getSolution().col_value)
        info = SimpleNamespace(status='solved' if ok else 'unsolved', iter=-1,
                               run_time=time.perf_counter() - t0,
                               obj_val=float(h.getInfo().objective_function_value))
        return SimpleNamespace(x=x, y=np.array(h.getSolution().row_dual), info=info)
