"""Oracle-A: literal restatement of the reference's autodiff path (TEST INFRA).

Same algorithm as the reference – per-sample value + forward-mode Jacobian,
vmapped over samples, dense packing, SciPy CSR -> CSC – with ``torch.func`` in
float64 on the CPU standing in for ``jax.vmap(jax.jacfwd(...))`` (JAX is not in
this image).  Every function cites the reference lines it follows.  Slow and
O(M^2) in memory by construction (that is what the reference does); usable for
M up to ~10^3.  See ``oracle/__init__.py`` for what pins parity.

JAX semantics that matter and that torch reproduces:
  d|v|/dv = sign(v) (0 at v = 0);  d||x||/dx = x/||x||;  C-order reshapes;
  ``mean`` over axis 0.
"""
import numpy as np
import scipy.sparse as sp
import torch
from torch.func import jacfwd, vmap

from riskaversetrajopt_b200.drone import drone_params as dp
from riskaversetrajopt_b200.car import driving_params as cp

_F64 = torch.float64


def _t(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=_F64)


# =============================================================================
# Quadrotor  (reference drone/drone_risk.py)
# =============================================================================
class DroneOracleA:
    """Restates ``Model`` of drone/drone_risk.py:70-469 (hot-path methods only).

    ``variant='risk'`` follows drone_risk.py; ``variant='times'`` applies the two
    differences of drone_times.py (relaxation 1e-5 / +-10 at :421-425, baseline
    rows without the 0.01 multiplier and without the -1e-3 padding at :324-334).
    """

    def __init__(self, S, DWs, masses, obs_Qs, method='saa', alpha=0.1,
                 variant='risk'):
        self.S, self.dt = int(S), dp.T / S             # :82-83
        self.method, self.alpha = method, alpha
        self.u_max, self.u_min = dp.u_max, -dp.u_max    # :85-86
        self.beta, self.drag = dp.beta, dp.drag_coefficient
        self.DWs, self.masses, self.obs_Qs = _t(DWs), _t(masses), _t(obs_Qs)
        self.M = self.masses.shape[0]
        self.variant = variant
        self._K = _t(dp.feedback_gain)
        self._x0, self._xf = _t(dp.x_init), _t(dp.x_final)
        self._obs_p = _t(dp.obs_positions)

    # -- drone_risk.py:108-120 (u_z left at 0) / drone_times.py:144 (all axes)
    def initial_guess_us_mat(self):
        us = np.zeros((self.S, dp.n_u))
        guess = (self.u_max + self.u_min) / 2.0 + 1e-2
        if self.variant == 'times':
            us[:, :] = guess
        else:
            us[:, :dp.n_u - 1] = guess
        return us

    # -- drone_risk.py:122-131
    def b(self, x, u, mass):
        v = x[3:6]
        applied = u + self._K @ x
        acc = applied / mass - self.drag * torch.abs(v) * v / mass
        return torch.cat([v, acc])

    # -- drone_risk.py:133-137
    def sigma(self, x, u, mass):
        s = torch.zeros((dp.n_x, dp.n_x), dtype=_F64)
        s[3:6, 3:6] = torch.eye(3, dtype=_F64)
        return s * (self.beta / mass)

    # -- drone_risk.py:139-155  (note the second sqrt(dt): dW is already scaled)
    def us_to_state_trajectory(self, us_mat, mass, dWs):
        x = self._x0
        out = [x]
        for t in range(self.S):
            drift = self.dt * self.b(x, us_mat[t], mass)
            diff = (self.dt ** 0.5) * (self.sigma(x, us_mat[t], mass) @ dWs[t])
            x = x + drift + diff
            out.append(x)
        return torch.stack(out)

    # -- drone_risk.py:157-162
    def us_to_state_trajectories(self, us_mat):
        us = _t(us_mat)
        Xs = vmap(self.us_to_state_trajectory, in_dims=(None, 0, 0))(
            us, self.masses, self.DWs)
        return Xs.numpy()

    # -- drone_risk.py:164-167
    def final_constraints(self, xs):
        return xs[-1] - self._xf

    # -- drone_risk.py:169-213  (g[o,k] for k = 1..S, uses xs[1:])
    def obstacle_avoidance_constraints(self, xs, obs_Q):
        rows = []
        for o in range(dp.n_obs):
            d = xs[1:, :2] - self._obs_p[o, :2]
            Q = obs_Q[o, :2, :2]
            rows.append(1.0 - torch.einsum('ka,ab,kb->k', d, Q, d))
        return torch.stack(rows)

    # -- drone_risk.py:239-280
    def get_all_constraints_coeffs(self, us_mat, mass, dWs, obs_Q):
        S = self.S

        def cons(u):
            xs = self.us_to_state_trajectory(u, mass, dWs)
            return self.final_constraints(xs), self.obstacle_avoidance_constraints(xs, obs_Q)

        v_final, g_obs = cons(us_mat)
        v_final_du, g_obs_du = jacfwd(cons)(us_mat)
        v_final_du = v_final_du.reshape(dp.n_x, dp.n_u * S)
        g_obs_du = g_obs_du.reshape(dp.n_obs, S, dp.n_u * S)
        u_vec = us_mat.reshape(S * dp.n_u)
        val_final = -v_final + v_final_du @ u_vec            # :271
        g_up = -g_obs + g_obs_du @ u_vec                     # :278
        return v_final_du, val_final, val_final, g_obs_du, g_up

    def per_sample(self, us_mat):
        """vmapped :288-290; returns numpy arrays."""
        us = _t(us_mat)
        out = vmap(self.get_all_constraints_coeffs, in_dims=(None, 0, 0, 0))(
            us, self.masses, self.DWs, self.obs_Qs)
        return tuple(o.numpy() for o in out)

    # -- drone_risk.py:282-374
    def get_all_constraints_coeffs_all(self, us_mat):
        S, M, n_obs = self.S, self.M, dp.n_obs
        nu = dp.n_u * S
        final_du, final_low, final_up, g_du, g_up = self.per_sample(us_mat)
        final_du = final_du.mean(axis=0)                     # :294-296
        final_low = final_low.mean(axis=0)
        final_up = final_up.mean(axis=0)
        final_dparams = np.concatenate(
            [final_du, np.zeros((final_du.shape[0], M + 2))], axis=-1)
        blk = n_obs * S
        if self.method == 'baseline':                        # :303-325
            mult = 1.0 if self.variant == 'times' else 0.01
            pad = 0.0 if self.variant == 'times' else 1e-3
            low = -np.inf * np.ones(M * blk)
            up = np.inf * np.ones(M * blk)
            D = np.zeros((M * blk, nu + M + 2))
            for i in range(M):
                r0, r1 = i * blk, (i + 1) * blk
                D[r0:r1, :nu] = (mult * g_du[i]).reshape(blk, nu)
                up[r0:r1] = mult * g_up[i].reshape(-1)
                up[r0:r1] = up[r0:r1] - pad
        else:                                                # :327-368
            nrow = 1 + M + M * blk + 1
            low = -np.inf * np.ones(nrow)
            up = np.inf * np.ones(nrow)
            D = np.zeros((nrow, nu + M + 2))
            D[0, -1] = M * self.alpha
            D[0, nu:-1] = 1.0          # covers y_0..y_{M-1} AND the slack column
            up[0] = 0.0
            mult = 0.01
            for i in range(M):
                col_y = nu + i
                D[1 + i, col_y] = -1.0
                up[1 + i] = 0.0
                D[1 + i, -2] = -1.0
                r0, r1 = 1 + M + i * blk, 1 + M + (i + 1) * blk
                D[r0:r1, :nu] = (mult * g_du[i]).reshape(blk, nu)
                D[r0:r1, col_y] = -mult
                up[r0:r1] = mult * g_up[i].reshape(-1)
                D[r0:r1, -1] = -mult
            D[-1, -2] = -1.0
            up[-1] = 0.0
        return (np.vstack([final_dparams, D]),
                np.hstack([final_low, low]),
                np.hstack([final_up, up]))

    # -- drone_risk.py:221-237
    def get_control_constraints_coeffs_all(self):
        nu = dp.n_u * self.S
        A = np.zeros((nu, nu + self.M + 2))
        A[np.arange(nu), np.arange(nu)] = 1.0
        return A, self.u_min * np.ones(nu), self.u_max * np.ones(nu)

    # -- drone_risk.py:376-399
    def get_objective_coeffs(self):
        S, M = self.S, self.M
        n = dp.n_u * S + M + 2
        P = np.zeros((n, n))
        q = np.zeros(n)
        for t in range(S):
            i = t * dp.n_u
            P[i:i + dp.n_u, i:i + dp.n_u] = 2 * self.dt * dp.R
        P[-2, -2] = 10000.0
        q[-2] = 10000.0
        return sp.csc_matrix(P), q

    # -- drone_risk.py:401-423  (the drop-in boundary)
    def get_constraints_coeffs(self, us_mat, scp_iter):
        A_con, l_con, u_con = self.get_control_constraints_coeffs_all()
        As, ls, us = self.get_all_constraints_coeffs_all(us_mat)
        As, ls, us = np.copy(As), np.copy(ls), np.copy(us)
        if scp_iter < 2:
            if self.variant == 'times':                      # drone_times.py:421-425
                As[dp.n_x:] *= 1e-5
                ls[dp.n_x:] = -10.0
                us[dp.n_x:] = 10.0
            else:                                            # drone_risk.py:413-417
                As[dp.n_x:] *= 1e-7
                ls[dp.n_x:] = -0.1
                us[dp.n_x:] = 0.1
        A = sp.vstack([sp.csr_matrix(As), sp.csr_matrix(A_con)], format='csc')
        return A, np.hstack([ls, l_con]), np.hstack([us, u_con])

    # -- drone_risk.py:656-662 (Monte-Carlo verification, per sample)
    def monte_carlo_constraints(self, us_mat):
        us = _t(us_mat)

        def one(mass, dWs, obs_Q):
            xs = self.us_to_state_trajectory(us, mass, dWs)
            return torch.max(self.obstacle_avoidance_constraints(xs, obs_Q)) - dp.OSQP_TOL

        Z = vmap(one)(self.masses, self.DWs, self.obs_Qs).numpy()
        return Z <= 1e-6, Z


def avar_closed_form(Z, t_risk, alpha):
    """drone_risk.py:694 / car/driving.py:670 / hopper/hopper.py:957."""
    Z = np.asarray(Z)
    return t_risk + np.mean(np.maximum(Z - t_risk, np.zeros(len(Z))) / alpha)


# =============================================================================
# Car + pedestrian  (reference car/driving.py)
# =============================================================================
def car_sample_parameters(M, method='saa'):
    """RNG call order of ``Model.__init__`` (car/driving.py:95-120), legacy
    global stream: M uniforms (omega_speed), M uniforms (omega_repulsive), for
    saa M x randn(4) scaled by sqrt(variance), then M*S*n_x normals * sqrt(dt).
    -> (states_init (M,8), omegas_speed (M,), omegas_repulsive (M,), DWs (M,S,8))
    """
    w_s = np.random.uniform(cp.omega_speed_nom - cp.omega_speed_del,
                            cp.omega_speed_nom + cp.omega_speed_del, M)
    w_r = np.random.uniform(cp.omega_repulsive_nom - cp.omega_repulsive_del,
                            cp.omega_repulsive_nom + cp.omega_repulsive_del, M)
    states_init = np.repeat(cp.state_init[None, :], M, axis=0)
    if method == 'saa':
        std = np.sqrt(cp.variance_ped_initial_state)
        for i in range(M):
            states_init[i, 4:] = states_init[i, 4:] + std @ np.random.randn(4)
    DWs = np.zeros((M, cp.S, cp.n_x))
    for i in range(M):
        for t in range(cp.S):
            DWs[i, t, :] = np.random.randn(cp.n_x)
    DWs = np.sqrt(cp.dt) * DWs
    if method == 'baseline':
        DWs, w_s, w_r = 0 * DWs, 0 * w_s, 0 * w_r
    return states_init, w_s, w_r, DWs


class CarOracleA:
    """Restates ``Model`` of car/driving.py:83-456 (hot-path methods only)."""

    def __init__(self, states_init, omegas_speed, omegas_repulsive, DWs,
                 method='saa', alpha=0.05):
        self.method, self.alpha = method, alpha
        self.u_max, self.u_min = cp.u_max, -cp.u_max
        self.beta = 3e-2                                      # :94
        self.states_init = _t(states_init)
        self.omegas_speed = _t(omegas_speed)
        self.omegas_repulsive = _t(omegas_repulsive)
        self.DWs = _t(DWs)
        self.M = self.omegas_speed.shape[0]
        self.S, self.dt = cp.S, cp.dt
        self._goal = _t(np.concatenate([cp.position_ego_goal, cp.velocity_ego_goal]))

    def initial_guess_us_mat(self):                           # :132-143
        return np.full((self.S, cp.n_u), (self.u_max + self.u_min) / 2.0 + 1e-2)

    def force_on_pedestrian(self, x, w_s, w_r):               # :146-158
        delta = x[0:2] - x[4:6]
        force = -w_r * delta
        force = force / torch.linalg.norm(delta)
        return force + w_s * (cp.speed_ped_des - x[7])        # scalar broadcast onto both

    def b(self, x, u, w_s, w_r):                              # :161-178
        F = self.force_on_pedestrian(x, w_s, w_r)
        return torch.stack([x[2] * torch.cos(x[3]), x[2] * torch.sin(x[3]),
                            u[0], u[1], x[6], x[7], F[0], F[1]])

    def sigma(self, x, u):                                    # :181-184
        s = torch.zeros((cp.n_x, cp.n_x), dtype=_F64)
        s[6, 6] = self.beta
        s[7, 7] = self.beta
        return s

    def us_to_state_trajectory(self, us_mat, state_init, w_s, w_r, dWs):  # :187-204
        x = state_init
        out = [x]
        for t in range(self.S):
            drift = self.dt * self.b(x, us_mat[t], w_s, w_r)
            diff = (self.dt ** 0.5) * (self.sigma(x, us_mat[t]) @ dWs[t])
            x = x + drift + diff
            out.append(x)
        return torch.stack(out)

    def us_to_state_trajectories(self, us_mat):               # :207-214
        us = _t(us_mat)
        return vmap(self.us_to_state_trajectory, in_dims=(None, 0, 0, 0, 0))(
            us, self.states_init, self.omegas_speed, self.omegas_repulsive,
            self.DWs).numpy()

    def final_constraints(self, xs):                          # :217-221
        return xs[-1, :4] - self._goal

    def separation_distances_at_all_times(self, xs):          # :224-236
        d = xs[1:, 0:2] - xs[1:, 4:6]
        return torch.linalg.norm(d, dim=-1) - float(cp.min_separation_distance)

    def get_all_constraints_coeffs(self, us_mat, state_init, w_s, w_r, dWs):  # :261-298
        S = self.S

        def cons(u):
            xs = self.us_to_state_trajectory(u, state_init, w_s, w_r, dWs)
            return self.final_constraints(xs), -self.separation_distances_at_all_times(xs)

        v_final, g_obs = cons(us_mat)
        v_final_du, g_obs_du = jacfwd(cons)(us_mat)
        v_final_du = v_final_du.reshape(4, cp.n_u * S)
        g_obs_du = g_obs_du.reshape(S, cp.n_u * S)
        u_vec = us_mat.reshape(S * cp.n_u)
        val_final = -v_final + v_final_du @ u_vec
        g_up = -g_obs + g_obs_du @ u_vec
        return v_final_du, val_final, val_final, g_obs_du, g_up

    def per_sample(self, us_mat):
        us = _t(us_mat)
        out = vmap(self.get_all_constraints_coeffs, in_dims=(None, 0, 0, 0, 0))(
            us, self.states_init, self.omegas_speed, self.omegas_repulsive, self.DWs)
        return tuple(o.numpy() for o in out)

    def get_all_constraints_coeffs_all(self, us_mat):         # :302-373
        S, M = self.S, self.M
        nu = cp.n_u * S
        final_du, final_low, final_up, g_du, g_up = self.per_sample(us_mat)
        final_du = final_du.mean(axis=0)
        final_low = final_low.mean(axis=0)
        final_up = final_up.mean(axis=0)
        final_dparams = np.concatenate(
            [final_du, np.zeros((final_du.shape[0], M + 2))], axis=-1)
        if self.method == 'baseline':                         # :321-330
            low = -np.inf * np.ones(M * S)
            up = np.inf * np.ones(M * S)
            D = np.zeros((M * S, nu + M + 2))
            for i in range(M):
                D[i * S:(i + 1) * S, :nu] = g_du[i].reshape(S, nu)
                up[i * S:(i + 1) * S] = g_up[i].reshape(-1)
        else:                                                 # :332-367
            nrow = 1 + M + M * S + 1
            low = -np.inf * np.ones(nrow)
            up = np.inf * np.ones(nrow)
            D = np.zeros((nrow, nu + M + 2))
            D[0, -1] = M * self.alpha
            D[0, nu:-1] = 1.0
            up[0] = 0.0
            for i in range(M):
                col_y = nu + i
                D[1 + i, col_y] = -1.0
                up[1 + i] = 0.0
                D[1 + i, -2] = -1.0
                r0, r1 = 1 + M + i * S, 1 + M + (i + 1) * S
                D[r0:r1, :nu] = g_du[i].reshape(S, nu)
                D[r0:r1, col_y] = -1.0
                up[r0:r1] = g_up[i].reshape(-1)
                D[r0:r1, -1] = -1.0
            D[-1, -2] = -1.0
            up[-1] = 0.0
        return (np.vstack([final_dparams, D]),
                np.hstack([final_low, low]),
                np.hstack([final_up, up]))

    def get_control_constraints_coeffs_all(self):             # :243-258
        nu = cp.n_u * self.S
        A = np.zeros((nu, nu + self.M + 2))
        A[np.arange(nu), np.arange(nu)] = 1.0
        return A, self.u_min * np.ones(nu), self.u_max * np.ones(nu)

    def get_objective_coeffs(self):                           # :375-397
        S, M = self.S, self.M
        n = cp.n_u * S + M + 2
        P = np.zeros((n, n))
        q = np.zeros(n)
        for t in range(S):
            i = t * cp.n_u
            P[i:i + cp.n_u, i:i + cp.n_u] = 2 * self.dt * cp.R
        P[-2, -2] = 1000.0
        q[-2] = 1000.0
        return sp.csc_matrix(P), q

    def get_constraints_coeffs(self, us_mat, scp_iter):       # :399-421
        A_con, l_con, u_con = self.get_control_constraints_coeffs_all()
        As, ls, us = self.get_all_constraints_coeffs_all(us_mat)
        As, ls, us = np.copy(As), np.copy(ls), np.copy(us)
        if scp_iter < 1:
            # n_x = 8 although there are only 4 final rows: rows 4..7 (CVaR row,
            # -y_0..-y_2) survive; -inf * 0 = nan in ``ls`` is the reference's
            # behaviour and is kept.
            with np.errstate(invalid='ignore'):
                As[cp.n_x:] *= 0
                ls[cp.n_x:] *= 0
                us[cp.n_x:] *= 0
        A = sp.vstack([sp.csr_matrix(As), sp.csr_matrix(A_con)], format='csc')
        return A, np.hstack([ls, l_con]), np.hstack([us, u_con])

    def monte_carlo_constraints(self, us_mat):                # :630-638
        us = _t(us_mat)

        def one(s0, w_s, w_r, dWs):
            xs = self.us_to_state_trajectory(us, s0, w_s, w_r, dWs)
            return torch.max(-self.separation_distances_at_all_times(xs)) - cp.OSQP_TOL

        Z = vmap(one)(self.states_init, self.omegas_speed,
                      self.omegas_repulsive, self.DWs).numpy()
        return Z <= 1e-6, Z
