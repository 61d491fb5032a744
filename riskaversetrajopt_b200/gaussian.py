"""Gaussian-linearisation baselines of the reference, as a HOST comparator (SURVEY.md 8f row 4,
BASELINE config 5: "SAA vs Gaussian-approx constraints at M = 10^6").

The reference's ``drone/drone_gaussian.py`` and ``car/driving_gaussian.py`` replace the samples by
one nominal trajectory and a covariance recursion (Lew, Bonalli, Pavone, ECC 2020)

    Sigma_{t+1} = A_t Sigma_t A_t^T + dt sigma sigma^T + sum_theta Var(theta) |dt b_theta|^2 (*),
    A_t = I + dt db/dx(x_t, u_t),
    (*) the paper's term is the outer product (dt b_theta)(dt b_theta)^T; the reference's code forms
        ``b @ b.T`` on 1-D arrays, i.e. the scalar |b|^2 broadcast onto every entry -- reproduced,

and tighten each constraint by a quantile of its linearised standard deviation.  There is no sample
axis (one 6x6 / 8x8 recursion over S = 20 steps), hence no GPU work: this module restates the
recursion and the chance-constraint values with closed-form Jacobians in NumPy, so that a control
sequence can be scored under BOTH models -- Gaussian constraint values here, sampled CVaR terms
on the device (``Model.monte_carlo_constraints`` at M = 10^6).  It is checked against the
reference's own code executed through ``oracle/refexec`` (tests/test_gaussian_cpu.py).
"""
import numpy as np
from scipy.stats import norm

from .drone import drone_params as dp
from .car import driving_params as cp


class DroneGaussian:
    """Nominal trajectory, covariance trajectory and chance constraints of
    drone/drone_gaussian.py:136-330 (same method names)."""

    def __init__(self, S=dp.S, alpha=0.1):
        self.S, self.dt, self.alpha = int(S), dp.T / S, alpha
        self.beta, self.drag = dp.beta, dp.drag_coefficient
        self.mass_nominal = dp.mass_nom
        self.mass_variance = (2 * dp.mass_delta) ** 2 / 12.0                 # :82 (uniform distribution)
        self.K = np.asarray(dp.feedback_gain, dtype=np.float64)
        self.obs_positions = np.asarray(dp.obs_positions, dtype=np.float64)
        self.obs_radii = np.asarray(dp.obs_radii, dtype=np.float64)

    def b(self, x, u, mass):                                                  # :136-145
        v = x[3:6]
        return np.concatenate([v, (u + self.K @ x) / mass - self.drag * np.abs(v) * v / mass])

    def b_dx(self, x, u):                                                     # :147-148 (jacfwd there)
        m, v = self.mass_nominal, x[3:6]
        J = np.zeros((6, 6))
        J[:3, 3:] = np.eye(3)
        J[3:, :] = self.K / m
        J[3:, 3:] += -np.diag(2.0 * self.drag * np.abs(v)) / m               # d(|v| v)/dv = 2|v|
        return J

    def b_dmass(self, x, u):                                                  # :151-152
        m, v = self.mass_nominal, x[3:6]
        acc = (u + self.K @ x) - self.drag * np.abs(v) * v
        return np.concatenate([np.zeros(3), -acc / m ** 2])

    def sigma(self):                                                          # :155-160
        s = np.zeros((6, 6))
        s[3:, 3:] = (self.beta / self.mass_nominal) * np.eye(3)
        return s

    def us_to_state_trajectory(self, us_mat):                                 # :162-175
        xs = np.zeros((self.S + 1, 6))
        xs[0] = np.asarray(dp.x_init, dtype=np.float64)
        for t in range(self.S):
            xs[t + 1] = xs[t] + self.dt * self.b(xs[t], us_mat[t], self.mass_nominal)
        return xs

    def us_to_covariance_trajectory(self, us_mat):                            # :177-227
        xs = self.us_to_state_trajectory(us_mat)
        Sig = np.zeros((self.S + 1, 6, 6))
        Sw = self.sigma()
        Sw = self.dt * Sw @ Sw.T
        for t in range(self.S):
            A = np.eye(6) + self.dt * self.b_dx(xs[t], us_mat[t])
            bm = self.dt * self.b_dmass(xs[t], us_mat[t])
            # Reference quirk, kept: b_dm is a 1-D array there, so ``b_dm @ b_dm.T`` (:209) is the
            # scalar |b_dm|^2 and is ADDED TO EVERY ENTRY of Sigma (broadcast), not the outer product
            Sig[t + 1] = A @ Sig[t] @ A.T + Sw + self.mass_variance * float(bm @ bm)
        return Sig

    def obstacle_avoidance_constraints(self, xs, Sigmas, alphas_risk_state, alphas_risk_obs):
        """g[o, k] <= 0 for k = 1..S (:239-318): -(||p - c_o|| - q(1 - alpha_ok) sqrt(n^T Sigma n) - r_o)
        with the obstacle radius at its alpha/3 quantile.  alphas_risk_state (S, n_obs), alphas_risk_obs (n_obs)."""
        n_obs = self.obs_positions.shape[0]
        g = np.zeros((n_obs, self.S))
        for o in range(n_obs):
            rad_min, rad_max = self.obs_radii[o] - dp.obs_radii_deltas, self.obs_radii[o] + dp.obs_radii_deltas
            radius = rad_max - (alphas_risk_obs[o] / 3.0) * (rad_max - rad_min)
            for k in range(1, self.S + 1):
                d = xs[k, :2] - self.obs_positions[o, :2]
                dist = np.linalg.norm(d)
                n = d / dist
                pad = norm.ppf(1 - alphas_risk_state[k - 1, o]) * np.sqrt(n @ Sigmas[k, :2, :2] @ n)
                g[o, k - 1] = -(dist - pad - radius)
        return g

    def uniform_risk_allocation(self):                                        # :118-124
        a = self.alpha / (self.S * 3 + 3)
        return a * np.ones((self.S, 3)), a * np.ones(3)


class CarGaussian:
    """car/driving_gaussian.py:66-262."""

    def __init__(self, alpha=0.1):
        self.alpha, self.beta = alpha, 3e-2
        self.S, self.dt = cp.S, cp.dt
        self.w_s, self.w_r = cp.omega_speed_nom, cp.omega_repulsive_nom
        self.w_s_var = (2 * cp.omega_speed_del) ** 2 / 12.0                   # :80-83
        self.w_r_var = (2 * cp.omega_repulsive_del) ** 2 / 12.0
        self.x0 = np.asarray(cp.state_init, dtype=np.float64)
        self.Sigma0 = np.zeros((8, 8))
        self.Sigma0[4:, 4:] = np.asarray(cp.variance_ped_initial_state, dtype=np.float64)

    def b(self, x, u, w_s, w_r):                                              # :114-142
        d = x[0:2] - x[4:6]
        F = -w_r * d / np.linalg.norm(d) + w_s * (cp.speed_ped_des - x[7])
        return np.array([x[2] * np.cos(x[3]), x[2] * np.sin(x[3]), u[0], u[1], x[6], x[7], F[0], F[1]])

    def b_dx(self, x, u):                                                     # :150-153
        J = np.zeros((8, 8))
        v, phi = x[2], x[3]
        J[0, 2], J[0, 3] = np.cos(phi), -v * np.sin(phi)
        J[1, 2], J[1, 3] = np.sin(phi), v * np.cos(phi)
        J[4, 6] = J[5, 7] = 1.0
        d = x[0:2] - x[4:6]
        nrm = np.linalg.norm(d)
        G = -self.w_r * (np.eye(2) / nrm - np.outer(d, d) / nrm ** 3)         # dF/dp_ego
        J[6:8, 0:2] = G
        J[6:8, 4:6] = -G
        J[6:8, 7] += -self.w_s
        return J

    def b_domega_speed(self, x, u):                                           # :155-158
        out = np.zeros(8)
        out[6:8] = cp.speed_ped_des - x[7]
        return out

    def b_domega_repulsive(self, x, u):                                       # :160-163
        d = x[0:2] - x[4:6]
        out = np.zeros(8)
        out[6:8] = -d / np.linalg.norm(d)
        return out

    def us_to_state_trajectory(self, us_mat):                                 # :171-186
        xs = np.zeros((self.S + 1, 8))
        xs[0] = self.x0
        for t in range(self.S):
            xs[t + 1] = xs[t] + self.dt * self.b(xs[t], us_mat[t], self.w_s, self.w_r)
        return xs

    def us_to_covariance_trajectory(self, us_mat):                            # :189-228
        xs = self.us_to_state_trajectory(us_mat)
        Sig = np.zeros((self.S + 1, 8, 8))
        Sig[0] = self.Sigma0
        Sw = np.zeros((8, 8))
        Sw[6:, 6:] = self.dt * self.beta ** 2 * np.eye(2)
        for t in range(self.S):
            A = np.eye(8) + self.dt * self.b_dx(xs[t], us_mat[t])
            bs, br = self.dt * self.b_domega_speed(xs[t], us_mat[t]), self.dt * self.b_domega_repulsive(xs[t], us_mat[t])
            # same quirk as the drone (:213-215): 1-D ``b @ b.T`` is a scalar, broadcast onto all of Sigma
            Sig[t + 1] = A @ Sig[t] @ A.T + Sw + (self.w_s_var * float(bs @ bs) + self.w_r_var * float(br @ br))
        return Sig

    def separation_distances_at_all_times(self, mus, Sigmas, alphas_risk):    # :238-264
        out = np.zeros(self.S)
        for k in range(1, self.S + 1):
            d = mus[k, 0:2] - mus[k, 4:6]
            dist = np.linalg.norm(d)
            n = d / dist
            pad = norm.ppf(1 - alphas_risk[k - 1]) * np.sqrt(n @ Sigmas[k, 4:6, 4:6] @ n)
            out[k - 1] = dist - pad - cp.min_separation_distance
        return out


def score_controls_both_ways(model, us_mat, alpha=None):
    """BASELINE config 5 in one call: a drone control sequence under the Gaussian model (host) and
    under the samples of ``model`` (a ``drone.drone_risk.Model``; device).  -> dict."""
    alpha = model.alpha if alpha is None else alpha
    g = DroneGaussian(model.S, alpha)
    xs, Sig = g.us_to_state_trajectory(us_mat), g.us_to_covariance_trajectory(us_mat)
    a_state, a_obs = g.uniform_risk_allocation()
    gg = g.obstacle_avoidance_constraints(xs, Sig, a_state, a_obs)
    sat, Z = model.monte_carlo_constraints(us_mat)
    t = float(np.quantile(Z, 1 - alpha))
    return {"gaussian_max_constraint": float(gg.max()), "gaussian_feasible": bool(gg.max() <= 0),
            "gaussian_max_position_std": float(np.sqrt(np.max(np.linalg.eigvalsh(Sig[1:, :2, :2])))),
            "saa_fraction_satisfied": float(np.mean(sat)), "saa_var_alpha": t,
            "saa_avar_alpha": model.monte_carlo_avar(us_mat, t, alpha), "samples": int(Z.size)}
