"""Oracle-B: analytic closed-form restatement, NumPy, scalable (TEST INFRA).

Same quantities as Oracle-A (hence as the reference) but with the derivatives
written out by hand instead of autodiff, vectorised over samples, and assembled
through SciPy COO -> CSC from explicit (row, col, value) triplets – i.e. the
sparsity pattern here is *structural* and is produced by SciPy's own sort, not
by the product's closed-form pattern builder, so it checks that builder
independently.  Must agree with Oracle-A to <= 1e-12 (tests/test_oracle.py).

Derivations (reference lines are what is being differentiated):

* drone, per axis a (axes decouple: diagonal gain drone_params.py:14-19, per-axis
  drag drone_risk.py:129-130):
      p+ = p + dt v
      v+ = v + dt (u - 0.05 p - 0.25 v - c|v|v)/m + sqrt(dt) (beta/m) dW
  so  d(p+,v+)/d(p,v) = [[1, dt], [-0.05 dt/m, 1 - dt (0.25 + 2c|v|)/m]],
      d v+/du = dt/m,   g[o,k] = 1 - sum_{a<2} Q_o,aa (p_k^a - c_o^a)^2.
* car: see ``CarOracleB``.
"""
import numpy as np
import scipy.sparse as sp

from riskaversetrajopt_b200.drone import drone_params as dp
from riskaversetrajopt_b200.car import driving_params as cp


# =============================================================================
# Quadrotor
# =============================================================================
class DroneOracleB:
    def __init__(self, S, DWs, masses, obs_Qs, method='saa', alpha=0.1,
                 variant='risk'):
        self.S, self.dt = int(S), dp.T / S
        self.method, self.alpha, self.variant = method, alpha, variant
        self.DWs = np.asarray(DWs, dtype=np.float64)
        self.masses = np.asarray(masses, dtype=np.float64)
        self.obs_Qs = np.asarray(obs_Qs, dtype=np.float64)
        self.M = self.masses.shape[0]

    # -- drone_risk.py:139-155 restated per axis --------------------------------
    def rollout(self, us_mat):
        """-> Xs (M, S+1, 6) and a22 (M, S, 3) = d v_{k+1}/d v_k."""
        S, dt, m = self.S, self.dt, self.masses
        us = np.asarray(us_mat, dtype=np.float64)
        c = dp.drag_coefficient
        Xs = np.empty((self.M, S + 1, 6))
        a22 = np.empty((self.M, S, 3))
        Xs[:, 0, :] = dp.x_init
        for t in range(S):
            p, v = Xs[:, t, :3], Xs[:, t, 3:]
            applied = us[t][None, :] + p @ dp.feedback_gain[:, :3].T + v @ dp.feedback_gain[:, 3:].T
            acc = applied / m[:, None] - c * np.abs(v) * v / m[:, None]
            noise = np.sqrt(dt) * (dp.beta / m)[:, None] * self.DWs[:, t, 3:6]
            Xs[:, t + 1, :3] = p + dt * v
            Xs[:, t + 1, 3:] = v + dt * acc + noise
            a22[:, t, :] = 1.0 - dt * (0.25 + 2.0 * c * np.abs(v)) / m[:, None]
        return Xs, a22

    def sensitivities(self, a22):
        """-> sp, sv (M, 3, S, S+1): d p_k^a / d u_{j,a}, d v_k^a / d u_{j,a}."""
        S, dt, m = self.S, self.dt, self.masses
        sp_ = np.zeros((self.M, 3, S, S + 1))
        sv_ = np.zeros((self.M, 3, S, S + 1))
        a21 = (-0.05 * dt / m)[:, None]
        for j in range(S):
            sv_[:, :, j, j + 1] = (dt / m)[:, None]
            for k in range(j + 1, S):
                sp_[:, :, j, k + 1] = sp_[:, :, j, k] + dt * sv_[:, :, j, k]
                sv_[:, :, j, k + 1] = a21 * sp_[:, :, j, k] + a22[:, k, :] * sv_[:, :, j, k]
        return sp_, sv_

    def per_sample(self, us_mat):
        """Same five arrays as Oracle-A's ``per_sample`` (drone_risk.py:239-280)."""
        S, M = self.S, self.M
        us = np.asarray(us_mat, dtype=np.float64)
        u_vec = us.reshape(S * 3)
        Xs, a22 = self.rollout(us)
        sp_, sv_ = self.sensitivities(a22)
        final_du = np.zeros((M, 6, 3 * S))
        for a in range(3):
            final_du[:, a, a::3] = sp_[:, a, :, S]
            final_du[:, 3 + a, a::3] = sv_[:, a, :, S]
        v_final = Xs[:, S, :] - dp.x_final
        val_final = -v_final + final_du @ u_vec
        g = np.empty((M, dp.n_obs, S))
        g_du = np.zeros((M, dp.n_obs, S, 3 * S))
        for o in range(dp.n_obs):
            d = Xs[:, 1:, :2] - dp.obs_positions[o, :2]                  # (M,S,2)
            Qd = np.stack([self.obs_Qs[:, o, 0, 0], self.obs_Qs[:, o, 1, 1]], -1)
            g[:, o, :] = 1.0 - np.sum(Qd[:, None, :] * d * d, axis=-1)
            for a in range(2):
                coef = -2.0 * Qd[:, None, a] * d[:, :, a]                # (M,S) over k=1..S
                # g_du[i,o,k-1,j*3+a] = coef[i,k-1] * dp_k/du_j
                g_du[:, o, :, a::3] = coef[:, :, None] * np.transpose(sp_[:, a, :, 1:], (0, 2, 1))
        g_up = -g + g_du @ u_vec
        return final_du, val_final, val_final.copy(), g_du, g_up

    # -- structural triplets of the assembled matrix -----------------------------
    def get_constraints_coeffs(self, us_mat, scp_iter):
        """(A csc, l, u) as drone_risk.py:401-423 returns them, built sparse."""
        S, M, n_obs = self.S, self.M, dp.n_obs
        nu, blk = 3 * S, dp.n_obs * S
        final_du, final_low, final_up, g_du, g_up = self.per_sample(us_mat)
        fdu = final_du.mean(axis=0)
        rows, cols, vals = [], [], []

        def add(r, c, v):
            r, c, v = np.broadcast_arrays(np.asarray(r), np.asarray(c), np.asarray(v, dtype=np.float64))
            rows.append(r.ravel()); cols.append(c.ravel()); vals.append(v.ravel())

        # final rows: structural nonzeros only (own axis; p-rows lack j = S-1)
        for a in range(3):
            j = np.arange(S - 1)
            add(a, j * 3 + a, fdu[a, j * 3 + a])
            j = np.arange(S)
            add(3 + a, j * 3 + a, fdu[3 + a, j * 3 + a])
        if self.method == 'baseline':
            mult = 1.0 if self.variant == 'times' else 0.01
            pad = 0.0 if self.variant == 'times' else 1e-3
            nrow = 6 + M * blk
            low = -np.inf * np.ones(M * blk)
            up = (mult * g_up).reshape(M * blk) - pad
            r_obs0 = 6
        else:
            mult = 0.01
            nrow = 6 + 1 + M + M * blk + 1
            low = -np.inf * np.ones(1 + M + M * blk + 1)
            up = np.concatenate([[0.0], np.zeros(M), (mult * g_up).reshape(M * blk), [0.0]])
            r_obs0 = 6 + 1 + M
            i = np.arange(M)
            add(6, nu + M + 1, M * self.alpha)
            add(6, nu + np.arange(M + 1), 1.0)           # y_i and slack (quirk :337)
            add(7 + i, nu + i, -1.0)
            add(7 + i, nu + M, -1.0)
            rr = r_obs0 + (i[:, None] * blk + np.arange(blk)[None, :])
            add(rr, nu + i[:, None], -mult)
            add(rr, nu + M + 1, -mult)
            add(nrow - 1, nu + M, -1.0)
        # u-block: rows (i,o,k) x cols (j,a), a<2, j<=k-2
        i = np.arange(M)
        for o in range(n_obs):
            for k in range(2, S + 1):
                for a in range(2):
                    j = np.arange(k - 1)
                    add(r_obs0 + i[:, None] * blk + o * S + (k - 1), (j * 3 + a)[None, :],
                        mult * g_du[:, o, k - 1, :][:, j * 3 + a])
        ls = np.hstack([final_low.mean(axis=0), low])
        us_ = np.hstack([final_up.mean(axis=0), up])
        rows, cols, vals = map(np.concatenate, (rows, cols, vals))
        if scp_iter < 2:
            scale, bound = (1e-5, 10.0) if self.variant == 'times' else (1e-7, 0.1)
            vals = np.where(rows >= dp.n_x, vals * scale, vals)
            ls[dp.n_x:] = -bound
            us_[dp.n_x:] = bound
        c = np.arange(nu)
        rows = np.concatenate([rows, nrow + c]); cols = np.concatenate([cols, c])
        vals = np.concatenate([vals, np.ones(nu)])
        A = sp.coo_matrix((vals, (rows, cols)), shape=(nrow + nu, nu + M + 2)).tocsc()
        A.sort_indices()
        l = np.hstack([ls, -dp.u_max * np.ones(nu)])
        u = np.hstack([us_, dp.u_max * np.ones(nu)])
        return A, l, u

    def monte_carlo_constraints(self, us_mat):
        """Z_i = max_{o,k} g - OSQP_TOL, sat_i = Z_i <= 1e-6 (drone_risk.py:656-662)."""
        Xs, _ = self.rollout(us_mat)
        Z = np.full(self.M, -np.inf)
        for o in range(dp.n_obs):
            d = Xs[:, 1:, :2] - dp.obs_positions[o, :2]
            Qd = np.stack([self.obs_Qs[:, o, 0, 0], self.obs_Qs[:, o, 1, 1]], -1)
            Z = np.maximum(Z, np.max(1.0 - np.sum(Qd[:, None, :] * d * d, axis=-1), axis=1))
        Z = Z - dp.OSQP_TOL
        return Z <= 1e-6, Z


# =============================================================================
# Car + pedestrian
# =============================================================================
class CarOracleB:
    """Analytic restatement of car/driving.py:145-421.  Forward sensitivities with
    the full 8x8 step Jacobian  A_k = I + dt db/dx  (ego block
    d(px',py')/d(v,phi) = dt [[cos, -v sin], [sin, v cos]]; pedestrian block
    dF/dp_ego = G = -w_r (I - n n^T)/|d| = -dF/dp_ped, dF_{0,1}/dx7 = -w_s) and
    B = dt [e_2, e_3]:  X_{k+1} = A_k X_k + B_k with X_k = d x_k / d u (8 x 40).
    Deliberately the generic dense recursion, not the reduced one the CUDA
    kernel uses, so that the two check each other."""

    def __init__(self, states_init, omegas_speed, omegas_repulsive, DWs, method='saa', alpha=0.05):
        self.method, self.alpha = method, alpha
        self.x0 = np.asarray(states_init, dtype=np.float64)
        self.w_s = np.asarray(omegas_speed, dtype=np.float64)
        self.w_r = np.asarray(omegas_repulsive, dtype=np.float64)
        self.DWs = np.asarray(DWs, dtype=np.float64)
        self.M, self.S, self.dt, self.beta = self.w_s.shape[0], int(np.shape(DWs)[1]), cp.dt, 3e-2   # S: horizon of the noise array (cp.S = 20 in the reference)

    def rollout(self, us_mat, with_jac=False):
        S, dt, M = self.S, self.dt, self.M
        us = np.asarray(us_mat, dtype=np.float64)
        Xs = np.empty((M, S + 1, 8))
        Xs[:, 0] = self.x0
        J = np.zeros((M, S + 1, 8, 2 * S)) if with_jac else None
        for t in range(S):
            x = Xs[:, t]
            d = x[:, 0:2] - x[:, 4:6]
            n = np.linalg.norm(d, axis=1)
            F = -self.w_r[:, None] * d / n[:, None] + (self.w_s * (cp.speed_ped_des - x[:, 7]))[:, None]
            b = np.stack([x[:, 2] * np.cos(x[:, 3]), x[:, 2] * np.sin(x[:, 3]),
                          np.full(M, us[t, 0]), np.full(M, us[t, 1]),
                          x[:, 6], x[:, 7], F[:, 0], F[:, 1]], axis=1)
            noise = np.zeros((M, 8))
            noise[:, 6:] = np.sqrt(dt) * self.beta * self.DWs[:, t, 6:]
            Xs[:, t + 1] = x + dt * b + noise
            if with_jac:
                A = np.zeros((M, 8, 8))
                A[:, 0, 2], A[:, 0, 3] = np.cos(x[:, 3]), -x[:, 2] * np.sin(x[:, 3])
                A[:, 1, 2], A[:, 1, 3] = np.sin(x[:, 3]), x[:, 2] * np.cos(x[:, 3])
                A[:, 4, 6] = 1.0
                A[:, 5, 7] = 1.0
                nh = d / n[:, None]
                G = -(self.w_r / n)[:, None, None] * (np.eye(2)[None] - nh[:, :, None] * nh[:, None, :])
                A[:, 6:8, 0:2] = G
                A[:, 6:8, 4:6] = -G
                A[:, 6, 7] += -self.w_s
                A[:, 7, 7] += -self.w_s
                A = np.eye(8)[None] + dt * A
                J[:, t + 1] = A @ J[:, t]
                J[:, t + 1, 2, 2 * t] += dt
                J[:, t + 1, 3, 2 * t + 1] += dt
        return (Xs, J) if with_jac else Xs

    def per_sample(self, us_mat):
        S = self.S
        us = np.asarray(us_mat, dtype=np.float64)
        u_vec = us.reshape(2 * S)
        Xs, J = self.rollout(us, with_jac=True)
        goal = np.concatenate([cp.position_ego_goal, cp.velocity_ego_goal])
        final_du = J[:, S, :4, :]
        val_final = -(Xs[:, S, :4] - goal) + final_du @ u_vec
        d = Xs[:, 1:, 0:2] - Xs[:, 1:, 4:6]
        n = np.linalg.norm(d, axis=2)
        g = -(n - float(cp.min_separation_distance))
        nh = d / n[:, :, None]
        g_du = -(nh[:, :, 0, None] * (J[:, 1:, 0] - J[:, 1:, 4]) + nh[:, :, 1, None] * (J[:, 1:, 1] - J[:, 1:, 5]))
        g_up = -g + g_du @ u_vec
        return final_du, val_final, val_final.copy(), g_du, g_up, g

    def get_constraints_coeffs(self, us_mat, scp_iter):
        S, M = self.S, self.M
        nu = 2 * S
        final_du, final_low, final_up, g_du, g_up, _ = self.per_sample(us_mat)
        fdu = final_du.mean(axis=0)
        rows, cols, vals = [], [], []

        def add(r, c, v):
            r, c, v = np.broadcast_arrays(np.asarray(r), np.asarray(c), np.asarray(v, dtype=np.float64))
            rows.append(r.ravel()); cols.append(c.ravel()); vals.append(v.ravel())

        j = np.arange(S - 1)
        for cc in range(2):
            add(0, j * 2 + cc, fdu[0, j * 2 + cc])
            add(1, j * 2 + cc, fdu[1, j * 2 + cc])
        j = np.arange(S)
        add(2, j * 2, fdu[2, j * 2])
        add(3, j * 2 + 1, fdu[3, j * 2 + 1])
        i = np.arange(M)
        if self.method == 'baseline':
            nrow = 4 + M * S
            low = -np.inf * np.ones(M * S)
            up = g_up.reshape(M * S).copy()
            r0 = 4
        else:
            nrow = 4 + 1 + M + M * S + 1
            low = -np.inf * np.ones(1 + M + M * S + 1)
            up = np.concatenate([[0.0], np.zeros(M), g_up.reshape(M * S), [0.0]])
            r0 = 4 + 1 + M
            add(4, nu + M + 1, M * self.alpha)
            add(4, nu + np.arange(M + 1), 1.0)
            add(5 + i, nu + i, -1.0)
            add(5 + i, nu + M, -1.0)
            rr = r0 + (i[:, None] * S + np.arange(S)[None, :])
            add(rr, nu + i[:, None], -1.0)
            add(rr, nu + M + 1, -1.0)
            add(nrow - 1, nu + M, -1.0)
        for k in range(2, S + 1):
            for cc in range(2):
                j = np.arange(k - 1)
                add(r0 + i[:, None] * S + (k - 1), (j * 2 + cc)[None, :], g_du[:, k - 1, :][:, j * 2 + cc])
        ls = np.hstack([final_low.mean(axis=0), low])
        us_ = np.hstack([final_up.mean(axis=0), up])
        rows, cols, vals = map(np.concatenate, (rows, cols, vals))
        if scp_iter < 1:
            keep = rows < cp.n_x                # rows >= n_x vanish (multiplied by exactly 0)
            rows, cols, vals = rows[keep], cols[keep], vals[keep]
            with np.errstate(invalid='ignore'):
                ls[cp.n_x:] *= 0
                us_[cp.n_x:] *= 0
        c = np.arange(nu)
        rows = np.concatenate([rows, nrow + c]); cols = np.concatenate([cols, c])
        vals = np.concatenate([vals, np.ones(nu)])
        A = sp.coo_matrix((vals, (rows, cols)), shape=(nrow + nu, nu + M + 2)).tocsc()
        A.sort_indices()
        return (A, np.hstack([ls, -cp.u_max * np.ones(nu)]), np.hstack([us_, cp.u_max * np.ones(nu)]))

    def monte_carlo_constraints(self, us_mat):
        _, _, _, _, _, g = self.per_sample(us_mat)
        Z = g.max(axis=1) - cp.OSQP_TOL
        return Z <= 1e-6, Z
