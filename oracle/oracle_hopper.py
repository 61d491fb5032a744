"""Hopper slip-risk oracles (TEST INFRA): Oracle-A = literal restatement of
hopper/hopper.py:75-81, :166-171, :300-367 with torch autodiff standing in for
jax.jacrev / jax.hessian (:568-575); Oracle-B = closed-form NumPy.
Parity is unpinned by the reference (no stored outputs); see oracle/__init__.py."""
import numpy as np
import torch
from torch.func import hessian, jacrev

S, n_x, n_u = 30, 8, 4
time_jump, time_land = 10, 20
mu_nom = 0.10
_F64 = torch.float64


def _t(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=_F64)


class HopperOracleA:
    def __init__(self, M, method, alpha, intensities, thetas, taus):
        self.M, self.method, self.alpha = M, method, alpha
        k = 0.0 if method == 'baseline' else 1.0            # hopper.py:96-103
        self.I, self.th, self.ta = _t(intensities) * k, _t(thetas) * k, _t(taus) * k

    # hopper.py:114-134
    def _split(self, Z):
        nxs, nus = (S + 1) * n_x, S * n_u
        xs = Z[:nxs].reshape(S + 1, n_x)
        us = Z[nxs:nxs + nus].reshape(S, n_u)
        return xs, us, Z[nxs + nus:-2], Z[-2], Z[-1]

    # hopper.py:75-81
    @staticmethod
    def friction_at_px(px, I, th, ta):
        return mu_nom + torch.sum(I * torch.cos(th * px + ta))

    # hopper.py:300-367
    def slip_risk_constraints(self, Z):
        xs, us, ys, slack, t_risk = self._split(Z)
        ee_x = xs[:, 0] + xs[:, 3] * torch.sin(xs[:, 2])                  # :166-171
        ee_x = torch.cat([ee_x[:time_jump], ee_x[time_land:-1]])
        forces = torch.cat([us[:time_jump, 2:], us[time_land:, 2:]])
        nc = forces.shape[0]
        rows = []
        for i in range(self.M):
            mu = torch.stack([self.friction_at_px(ee_x[c], self.I[i], self.th[i], self.ta[i]) for c in range(nc)])
            rows.append(forces[:, 0] - mu * forces[:, 1])
        cons = torch.stack(rows)                                          # (M, nc)
        if self.method == 'baseline':
            return (cons - slack).reshape(-1)
        head = ((self.M * self.alpha) * t_risk + torch.sum(ys)).reshape(1)
        body = (cons - t_risk - ys[:, None] - slack).reshape(-1)
        return torch.cat([head, -ys, body, torch.zeros(1, dtype=_F64)])

    def g(self, Z):
        return self.slip_risk_constraints(_t(Z)).numpy()

    def jac(self, Z):
        return jacrev(self.slip_risk_constraints)(_t(Z)).numpy()

    def hess(self, Z, lam):
        lam = _t(lam)
        return hessian(lambda z: torch.dot(lam, self.slip_risk_constraints(z)))(_t(Z)).numpy()


class HopperOracleB:
    """Closed form: per contact c at step t, p = x0 + x3 sin x2,
    row (i,c) = f_x - mu_i(p) f_z - t - y_i - slack."""

    def __init__(self, M, method, alpha, intensities, thetas, taus):
        self.M, self.method, self.alpha = M, method, alpha
        k = 0.0 if method == 'baseline' else 1.0
        self.I, self.th, self.ta = (np.asarray(a, dtype=np.float64) * k for a in (intensities, thetas, taus))
        self.t = np.concatenate([np.arange(0, time_jump), np.arange(time_land, S)])

    def friction(self, px):
        arg = self.th[:, None, :] * px[None, :, None] + self.ta[:, None, :]     # (M, nc, F)
        I, th = self.I[:, None, :], self.th[:, None, :]
        return (mu_nom + np.sum(I * np.cos(arg), -1), -np.sum(I * th * np.sin(arg), -1),
                -np.sum(I * th * th * np.cos(arg), -1))

    def g(self, Z):
        Z = np.asarray(Z, dtype=np.float64)
        xs = Z[:(S + 1) * n_x].reshape(S + 1, n_x)
        us = Z[(S + 1) * n_x:(S + 1) * n_x + S * n_u].reshape(S, n_u)
        ys, slack, t_risk = Z[(S + 1) * n_x + S * n_u:-2], Z[-2], Z[-1]
        px = xs[self.t, 0] + xs[self.t, 3] * np.sin(xs[self.t, 2])
        mu, _, _ = self.friction(px)
        cons = us[self.t, 2][None] - mu * us[self.t, 3][None]
        if self.method == 'baseline':
            return (cons - slack).reshape(-1)
        return np.concatenate([[self.M * self.alpha * t_risk + ys.sum()], -ys,
                               (cons - t_risk - ys[:, None] - slack).reshape(-1), [0.0]])
