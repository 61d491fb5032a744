"""NumPy restatement of the structure-exploiting ADMM that ``riskaversetrajopt_b200/device_qp.py`` runs on
the device (TEST INFRASTRUCTURE: only tests/ may import this; the product path is the CUDA library).

The QP of one SCP iteration (reference drone/drone_risk.py:327-399, car/driving.py:331-397) has the
variables (u, y_1..y_M, slack, t) and an "arrow" constraint matrix: ``nu`` dense u-columns, one column per
y_i that touches only sample i's rows (and the CVaR row), and the slack / t columns.  The algorithm is the
OSQP iteration of ``riskaversetrajopt_b200.qp.OSQPLike`` (Ruiz equilibration, reduced KKT system
``(P + sigma I + A' R A) x = rhs``); what changes is HOW the reduced system is solved:

* the y-block of ``K = P + sigma I + A' R A`` is diagonal once the CVaR row is taken out as a rank-one
  term ``rho_c v v'`` (Sherman-Morrison),
* the remaining coupling is eliminated sample by sample (Schur complement on the nu + 2 dense variables).

Everything sample-sized is then one pass over the samples per ADMM iteration with a handful of
reductions -- which is what shards across GPUs with a (nu + 4)-double all-reduce per iteration.

This file keeps the samples as dense ``J[i]`` (R x nu) arrays; it exists to pin the algebra (it is checked
against ``OSQPLike`` iterate for iterate in tests/test_arrow_admm_cpu.py) and to check the CUDA kernels.
"""
from types import SimpleNamespace

import numpy as np

_INF = 1e20


class ArrowQP:
    """Structured view of (P, q, A, l, u).  Column order of the QP: u (nu), y (M), slack, t.
    Row order: final (n_fin), CVaR, -y_i rows (M), sample rows (M * R), slack row, control rows (nu)."""

    def __init__(self, P, q, A, l, u, nu, M, n_fin, R):
        A = A.tocsc()
        self.nu, self.M, self.n_fin, self.R = nu, M, n_fin, R
        self.nw = nu + 2
        n = nu + M + 2
        r_cvar, r_y0, r_s0 = n_fin, n_fin + 1, n_fin + 1 + M
        r_sl, r_c0 = r_s0 + M * R, r_s0 + M * R + 1
        Ad = A.toarray()
        self.F = Ad[:n_fin, :nu].copy()
        self.J = Ad[r_s0:r_sl, :nu].reshape(M, R, nu).copy()
        # constants of the y / slack / t columns, read from the matrix itself
        self.cvar_y, self.cvar_s, self.cvar_t = Ad[r_cvar, nu], Ad[r_cvar, nu + M], Ad[r_cvar, nu + M + 1]
        self.yd, self.ys = Ad[r_y0, nu], Ad[r_y0, nu + M]
        self.yr, self.tr = Ad[r_s0, nu], Ad[r_s0, nu + M + 1]
        self.sl = Ad[r_sl, nu + M]
        self.ctrl = np.diag(Ad[r_c0:, :nu]).copy()
        P = np.asarray(P.todense()) if hasattr(P, 'todense') else np.asarray(P)
        self.P_uu = P[:nu, :nu].copy()
        self.P_ss, self.P_tt = P[n - 2, n - 2], P[n - 1, n - 1]
        self.q_u, self.q_s, self.q_t = q[:nu].copy(), q[n - 2], q[n - 1]
        self.set_bounds(l, u)

    def set_bounds(self, l, u):
        l = np.maximum(np.where(np.isnan(l), -_INF, l), -_INF)
        u = np.minimum(np.where(np.isnan(u), _INF, u), _INF)
        self.l, self.u = l, u

    def split_rows(self, v):
        """row vector -> (final, cvar, yrow (M), sample (M, R), slack, ctrl)"""
        nf, M, R = self.n_fin, self.M, self.R
        return (v[:nf], v[nf], v[nf + 1:nf + 1 + M], v[nf + 1 + M:nf + 1 + M + M * R].reshape(M, R),
                v[nf + 1 + M + M * R], v[nf + 2 + M + M * R:])


class ArrowADMM:
    def __init__(self, qp, eps_abs=1e-3, eps_rel=1e-3, max_iter=20000, rho=0.1, sigma=1e-6, alpha=1.6,
                 scaling=10, adaptive_rho_interval=50, check_interval=10):
        self.qp = qp
        self.o = SimpleNamespace(eps_abs=eps_abs, eps_rel=eps_rel, max_iter=max_iter, sigma=sigma, alpha=alpha,
                                 scaling=scaling, adaptive_rho_interval=adaptive_rho_interval,
                                 check_interval=check_interval)
        self.rho = rho
        nu, M, R, nf = qp.nu, qp.M, qp.R, qp.n_fin
        self.xw, self.xy = np.zeros(nu + 2), np.zeros(M)           # (u, slack, t), y   -- scaled variables
        m = nf + 1 + M + M * R + 1 + nu
        self.z, self.lam = np.zeros(m), np.zeros(m)
        self._scale()
        self._factor()

    # -- Ruiz equilibration, the same sequence of operations as OSQPLike._scale --------------------
    def _scale(self):
        qp, nu, M = self.qp, self.qp.nu, self.qp.M
        guard = lambda v: np.minimum(np.where(v < 1e-4, 1.0, v), 1e4)
        Du, Dy, Ds, Dt = np.ones(nu), np.ones(M), 1.0, 1.0
        EF, Ec, Ey, Es, Esl, Ect = np.ones(qp.n_fin), 1.0, np.ones(M), np.ones((M, qp.R)), 1.0, np.ones(nu)
        c = 1.0
        for _ in range(self.o.scaling):
            aJ = np.abs(qp.J) * Es[:, :, None] * Du[None, None, :]
            colP_u = c * np.max(np.abs(qp.P_uu) * Du[:, None] * Du[None, :], axis=0)
            colA_u = np.maximum.reduce([np.max(np.abs(qp.F) * EF[:, None] * Du[None, :], axis=0, initial=0.0),
                                        aJ.max(axis=(0, 1), initial=0.0), np.abs(qp.ctrl) * Ect * Du])
            colA_y = np.maximum.reduce([np.abs(qp.cvar_y) * Ec * Dy, np.abs(qp.yd) * Ey * Dy,
                                        np.abs(qp.yr) * Es.max(axis=1) * Dy])
            colA_s = max(abs(qp.cvar_s) * Ec * Ds, np.max(abs(qp.ys) * Ey * Ds, initial=0.0), abs(qp.sl) * Esl * Ds)
            colA_t = max(abs(qp.cvar_t) * Ec * Dt, abs(qp.tr) * Es.max(initial=0.0) * Dt)
            rowF = np.max(np.abs(qp.F) * EF[:, None] * Du[None, :], axis=1, initial=0.0)
            rowc = max(abs(qp.cvar_y) * Ec * Dy.max(initial=0.0), abs(qp.cvar_s) * Ec * Ds, abs(qp.cvar_t) * Ec * Dt)
            rowy = np.maximum(np.abs(qp.yd) * Ey * Dy, np.abs(qp.ys) * Ey * Ds)
            rows = np.maximum.reduce([aJ.max(axis=2, initial=0.0), np.abs(qp.yr) * Es * Dy[:, None],
                                      np.abs(qp.tr) * Es * Dt])
            rowsl = abs(qp.sl) * Esl * Ds
            rowct = np.abs(qp.ctrl) * Ect * Du
            du = 1.0 / np.sqrt(guard(np.maximum(colP_u, colA_u)))
            dy = 1.0 / np.sqrt(guard(colA_y))
            ds = 1.0 / np.sqrt(guard(np.array([max(c * abs(qp.P_ss) * Ds * Ds, colA_s)])))[0]
            dt = 1.0 / np.sqrt(guard(np.array([max(c * abs(qp.P_tt) * Dt * Dt, colA_t)])))[0]
            Du, Dy, Ds, Dt = Du * du, Dy * dy, Ds * ds, Dt * dt
            EF = EF / np.sqrt(guard(rowF)); Ec = Ec / np.sqrt(guard(np.array([rowc])))[0]
            Ey = Ey / np.sqrt(guard(rowy)); Es = Es / np.sqrt(guard(rows))
            Esl = Esl / np.sqrt(guard(np.array([rowsl])))[0]; Ect = Ect / np.sqrt(guard(rowct))
            # cost scaling
            colP = np.concatenate([c * np.max(np.abs(qp.P_uu) * Du[:, None] * Du[None, :], axis=0), np.zeros(M),
                                   [c * abs(qp.P_ss) * Ds * Ds, c * abs(qp.P_tt) * Dt * Dt]])
            qinf = c * max(np.max(np.abs(qp.q_u * Du), initial=0.0), abs(qp.q_s * Ds), abs(qp.q_t * Dt))
            g = 1.0 / max(np.mean(colP), qinf, 1e-4)
            c *= min(max(g, 1e-4), 1e4)
        self.Du, self.Dy, self.Ds, self.Dt = Du, Dy, Ds, Dt
        self.EF, self.Ec, self.Ey, self.Es, self.Esl, self.Ect = EF, Ec, Ey, Es, Esl, Ect
        self.c = c

    def _scaled_bounds(self):
        E = np.concatenate([self.EF, [self.Ec], self.Ey, self.Es.ravel(), [self.Esl], self.Ect])
        ls, us = E * self.qp.l, E * self.qp.u
        ls[self.qp.l <= -_INF], us[self.qp.u >= _INF] = -_INF, _INF
        return E, ls, us

    def _rho_rows(self):
        E, ls, us = self._scaled_bounds()
        r = np.full(ls.shape, self.rho)
        r[np.abs(us - ls) < 1e-10] = 1e3 * self.rho
        r[(ls <= -_INF) & (us >= _INF)] = 1e-6
        return r

    # -- scaled pieces ---------------------------------------------------------------------
    def _pieces(self):
        qp = self.qp
        Js = qp.J * self.Es[:, :, None] * self.Du[None, None, :]            # E J D_u
        Fs = qp.F * self.EF[:, None] * self.Du[None, :]
        e = self.Ec * qp.cvar_y * self.Dy                                   # CVaR row over y
        vcw = np.concatenate([np.zeros(qp.nu), [self.Ec * qp.cvar_s * self.Ds, self.Ec * qp.cvar_t * self.Dt]])
        ydS, ysS = self.Ey * qp.yd * self.Dy, self.Ey * qp.ys * self.Ds     # -y_i rows
        yrS, trS = self.Es * qp.yr * self.Dy[:, None], self.Es * qp.tr * self.Dt     # sample rows: y_i and t coefficients
        slS = self.Esl * qp.sl * self.Ds
        ctS = self.Ect * qp.ctrl * self.Du
        return Js, Fs, e, vcw, ydS, ysS, yrS, trS, slS, ctS

    def _factor(self):
        """Schur complement of K0 = P + sigma I + A0' R A0 (A0: all rows but the CVaR row) on the dense
        variables w = (u, slack, t), and the Sherman-Morrison vectors of the CVaR row."""
        qp, nu, M, sg = self.qp, self.qp.nu, self.qp.M, self.o.sigma
        Js, Fs, e, vcw, ydS, ysS, yrS, trS, slS, ctS = self._pieces()
        rr = self._rho_rows()
        rF, rc, ry, rs, rsl, rct = qp.split_rows(rr)
        self.rF, self.rc, self.ry, self.rs, self.rsl, self.rct = rF, rc, ry, rs, rsl, rct
        nw = nu + 2
        Kww = np.zeros((nw, nw))
        Kww[:nu, :nu] = self.c * qp.P_uu * self.Du[:, None] * self.Du[None, :]
        Kww[nu, nu] = self.c * qp.P_ss * self.Ds ** 2
        Kww[nu + 1, nu + 1] = self.c * qp.P_tt * self.Dt ** 2
        Kww += sg * np.eye(nw)
        Kww[:nu, :nu] += Fs.T @ (rF[:, None] * Fs) + np.diag(rct * ctS ** 2)
        # sample rows: g_ir = [Js_ir, 0, trS_ir]
        Kww[:nu, :nu] += np.einsum('irc,ir,ird->cd', Js, rs, Js)
        ut = np.einsum('irc,ir,ir->c', Js, rs, trS)
        Kww[:nu, nu + 1] += ut; Kww[nu + 1, :nu] += ut
        Kww[nu + 1, nu + 1] += np.sum(rs * trS ** 2)
        Kww[nu, nu] += np.sum(ry * ysS ** 2) + rsl * slS ** 2
        # y block (diagonal) and coupling b_i
        a = sg + ry * ydS ** 2 + np.sum(rs * yrS ** 2, axis=1)
        B = np.zeros((M, nw))
        B[:, :nu] = np.einsum('irc,ir,ir->ic', Js, rs, yrS)
        B[:, nu] = ry * ydS * ysS
        B[:, nu + 1] = np.sum(rs * yrS * trS, axis=1)
        S = Kww - B.T @ (B / a[:, None])
        self.a, self.B = a, B
        self.Sinv = np.linalg.inv(S)
        self.h = B.T @ (e / a)
        self.eps_c = float(np.sum(e * e / a))
        self.pw = self.Sinv @ (vcw - self.h)
        self.vp = float(vcw @ self.pw + self.eps_c - self.h @ self.pw)          # v_c' K0^{-1} v_c

    # -- one ADMM iteration: the operations of OSQPLike.solve, structured ------------------------------
    def _rhs(self):
        """r = sigma x - q + As'(rho z - lam), split into (r_w, r_y local part, gamma)."""
        qp, nu = self.qp, self.qp.nu
        Js, Fs, e, vcw, ydS, ysS, yrS, trS, slS, ctS = self._pieces()
        rr = np.concatenate([self.rF, [self.rc], self.ry, self.rs.ravel(), [self.rsl], self.rct])
        wF, wc, wy, ws, wsl, wct = qp.split_rows(rr * self.z - self.lam)
        qw = self.c * np.concatenate([qp.q_u * self.Du, [qp.q_s * self.Ds, qp.q_t * self.Dt]])
        rw = self.o.sigma * self.xw - qw
        rw[:nu] += Fs.T @ wF + np.einsum('irc,ir->c', Js, ws) + ctS * wct
        rw[nu] += np.sum(ysS * wy) + slS * wsl
        rw[nu + 1] += np.sum(trS * ws)
        ry_loc = self.o.sigma * self.xy + ydS * wy + np.sum(yrS * ws, axis=1)
        return rw, ry_loc, wc, vcw, e

    def _solve_K(self, rw, ry_loc, gamma, vcw, e):
        """x = K^{-1} r with r = (rw + gamma vcw, ry_loc + gamma e)."""
        a, B = self.a, self.B
        x0w = self.Sinv @ (rw + gamma * vcw - B.T @ (ry_loc / a) - gamma * self.h)
        vx0 = vcw @ x0w + np.sum(e * ry_loc / a) + gamma * self.eps_c - self.h @ x0w
        kappa = self.rc * vx0 / (1.0 + self.rc * self.vp)
        xw = x0w - kappa * self.pw
        xy = (ry_loc + e * (gamma - kappa) - B @ xw) / a
        return xw, xy

    def _Ax(self, xw, xy):
        qp, nu = self.qp, self.qp.nu
        Js, Fs, e, vcw, ydS, ysS, yrS, trS, slS, ctS = self._pieces()
        u, s, t = xw[:nu], xw[nu], xw[nu + 1]
        return np.concatenate([Fs @ u, [vcw @ xw + e @ xy], ydS * xy + ysS * s,
                               (np.einsum('irc,c->ir', Js, u) + yrS * xy[:, None] + trS * t).ravel(),
                               [slS * s], ctS * u])

    def _Aty(self, lam):
        qp, nu = self.qp, self.qp.nu
        Js, Fs, e, vcw, ydS, ysS, yrS, trS, slS, ctS = self._pieces()
        wF, wc, wy, ws, wsl, wct = qp.split_rows(lam)
        gw = wc * vcw
        gw[:nu] += Fs.T @ wF + np.einsum('irc,ir->c', Js, ws) + ctS * wct
        gw[nu] += np.sum(ysS * wy) + slS * wsl
        gw[nu + 1] += np.sum(trS * ws)
        gy = wc * e + ydS * wy + np.sum(yrS * ws, axis=1)
        return gw, gy

    def solve(self):
        o, qp, nu = self.o, self.qp, self.qp.nu
        E, ls, us = self._scaled_bounds()
        rr = np.concatenate([self.rF, [self.rc], self.ry, self.rs.ravel(), [self.rsl], self.rct])
        status, it = 'maximum iterations reached', 0
        for it in range(1, o.max_iter + 1):
            rw, ry_loc, gamma, vcw, e = self._rhs()
            xtw, xty = self._solve_K(rw, ry_loc, gamma, vcw, e)
            zt = self._Ax(xtw, xty)
            self.xw = o.alpha * xtw + (1 - o.alpha) * self.xw
            self.xy = o.alpha * xty + (1 - o.alpha) * self.xy
            zr = o.alpha * zt + (1 - o.alpha) * self.z
            z_new = np.clip(zr + self.lam / rr, ls, us)
            self.lam = self.lam + rr * (zr - z_new)
            self.z = z_new
            if it % o.check_interval == 0 or it == o.max_iter:
                rp, rd, ep, ed = self._residuals(E)
                if rp <= ep and rd <= ed:
                    status = 'solved'
                    break
                if o.adaptive_rho_interval and it % o.adaptive_rho_interval == 0:
                    num, den = rp / max(ep, 1e-30), rd / max(ed, 1e-30)
                    new_rho = float(np.clip(self.rho * np.sqrt(num / max(den, 1e-30)), 1e-6, 1e6))
                    if new_rho > 5 * self.rho or new_rho < self.rho / 5:
                        self.rho = new_rho
                        self._factor()
                        rr = np.concatenate([self.rF, [self.rc], self.ry, self.rs.ravel(), [self.rsl], self.rct])
        x = np.concatenate([self.Du * self.xw[:nu], self.Dy * self.xy, [self.Ds * self.xw[nu], self.Dt * self.xw[nu + 1]]])
        return SimpleNamespace(x=x, y=E * self.lam / self.c, info=SimpleNamespace(status=status, iter=it))

    def _residuals(self, E):
        o, qp, nu = self.o, self.qp, self.qp.nu
        Ax = self._Ax(self.xw, self.xy)
        gw, gy = self._Aty(self.lam)
        Pw = np.concatenate([self.c * (qp.P_uu * self.Du[:, None] * self.Du[None, :]) @ self.xw[:nu],
                             [self.c * qp.P_ss * self.Ds ** 2 * self.xw[nu], self.c * qp.P_tt * self.Dt ** 2 * self.xw[nu + 1]]])
        qw = self.c * np.concatenate([qp.q_u * self.Du, [qp.q_s * self.Ds, qp.q_t * self.Dt]])
        Dw = np.concatenate([self.Du, [self.Ds, self.Dt]])
        rp = np.max(np.abs((Ax - self.z) / E))
        rd = max(np.max(np.abs((Pw + qw + gw) / Dw)), np.max(np.abs(gy / self.Dy), initial=0.0)) / self.c
        ep = o.eps_abs + o.eps_rel * max(np.max(np.abs(Ax / E)), np.max(np.abs(self.z / E)))
        ed = o.eps_abs + o.eps_rel * max(np.max(np.abs(Pw / Dw)), np.max(np.abs(gw / Dw)),
                                         np.max(np.abs(gy / self.Dy), initial=0.0), np.max(np.abs(qw / Dw))) / self.c
        return rp, rd, ep, ed

    def update(self, A=None, l=None, u=None):
        """New matrix values / bounds (same scaling, as OSQP's update)."""
        if A is not None:
            q = self.qp
            new = ArrowQP(_FakeP(q), np.concatenate([q.q_u, np.zeros(q.M), [q.q_s, q.q_t]]), A,
                          q.l if l is None else l, q.u if u is None else u, q.nu, q.M, q.n_fin, q.R)
            self.qp = new
        elif l is not None or u is not None:
            self.qp.set_bounds(self.qp.l if l is None else l, self.qp.u if u is None else u)
        self._factor()


class _FakeP:
    def __init__(self, q):
        self.q = q

    def todense(self):
        q = self.q
        n = q.nu + q.M + 2
        P = np.zeros((n, n))
        P[:q.nu, :q.nu] = q.P_uu
        P[n - 2, n - 2], P[n - 1, n - 1] = q.P_ss, q.P_tt
        return P
