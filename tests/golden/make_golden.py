"""Mint the golden vectors under tests/golden/ from Oracle-A (the literal
autodiff restatement of the reference) at the reference's seeds.

    python tests/golden/make_golden.py

The reference itself cannot run in this image (no jax/osqp/ipyopt), ships no
stored outputs and has no tests, so these files are minted from the oracle and
pin it against regressions; what pins the oracle is described in
oracle/__init__.py.  Inputs are regenerated from the seeds at test time, only
outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.oracle_a import DroneOracleA  # noqa: E402
from riskaversetrajopt_b200.drone import drone_params as dp  # noqa: E402
from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters  # noqa: E402


def drone():
    np.random.seed(0)                                   # drone_risk.py:57
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=dp.M)   # first repeat, :483-484
    # (1) reference defaults: M = 50, alpha = 0.1, initial guess, scp_iter = 2
    m = DroneOracleA(dp.S, DWs, masses, obs_Qs, 'saa', 0.1)
    us = m.initial_guess_us_mat()
    A, l, u = m.get_constraints_coeffs(us, 2)
    Xs = m.us_to_state_trajectories(us)
    sat, Z = m.monte_carlo_constraints(us)
    np.savez_compressed(os.path.join(HERE, "drone_M50_saa_iter2.npz"), us=us, shape=A.shape,
                        indptr=A.indptr, indices=A.indices, data=A.data, l=l, u=u,
                        Xs_first3=Xs[:3], Z=Z)
    # (2) first 8 samples, perturbed iterate, all branches
    rs = np.random.RandomState(123)
    us = m.initial_guess_us_mat() + 0.5 * rs.randn(dp.S, dp.n_u)
    out = dict(us=us)
    for method in ('saa', 'baseline'):
        for variant in ('risk', 'times'):
            mm = DroneOracleA(dp.S, DWs[:8], masses[:8], obs_Qs[:8], method, 0.05, variant)
            for it in (0, 2):
                A, l, u = mm.get_constraints_coeffs(us, it)
                k = f"{method}_{variant}_{it}"
                out[k + "_indptr"], out[k + "_indices"] = A.indptr, A.indices
                out[k + "_data"], out[k + "_l"], out[k + "_u"] = A.data, l, u
    np.savez_compressed(os.path.join(HERE, "drone_M8_branches.npz"), **out)


def car():
    from oracle.oracle_a import CarOracleA, car_sample_parameters
    np.random.seed(0)                                   # car/driving.py:61
    samples = car_sample_parameters(50, 'saa')          # first Model(M, 'saa', alpha) of :472
    m = CarOracleA(*samples, 'saa', 0.05)
    us0 = m.initial_guess_us_mat()
    us1 = us0 + 0.4 * np.random.RandomState(7).randn(20, 2)
    out = dict(us0=us0, us1=us1)
    for name, us, it in (("iter0", us0, 0), ("iter1", us1, 1), ("iter2", us1, 2)):
        A, l, u = m.get_constraints_coeffs(us, it)
        out[name + "_shape"] = np.array(A.shape)
        out[name + "_indptr"], out[name + "_indices"], out[name + "_data"] = A.indptr, A.indices, A.data
        out[name + "_l"], out[name + "_u"] = l, u
    out["Xs_first3"] = m.us_to_state_trajectories(us1)[:3]
    out["Z"] = m.monte_carlo_constraints(us1)[1]
    np.savez_compressed(os.path.join(HERE, "car_M50_saa.npz"), **out)


def hopper():
    from oracle.oracle_hopper import HopperOracleA
    from riskaversetrajopt_b200.hopper import hopper as hp     # host-only constants + sampler
    np.random.seed(1)                                   # hopper/hopper.py:33
    feats = hp.sample_friction_features(hp.M)           # :70-74 (I, theta, tau), M = 30
    rs = np.random.RandomState(11)
    nv = hp.num_vars(hp.M)
    Z = np.zeros(nv)
    # a point inside the reference's variable box (:598-620)
    xs = np.zeros((hp.S + 1, 8))
    xs[:, 0] = rs.uniform(-0.2, 0.4, hp.S + 1); xs[:, 1] = rs.uniform(0.6, 1.2, hp.S + 1)
    xs[:, 2] = rs.uniform(-0.5, 0.5, hp.S + 1); xs[:, 3] = rs.uniform(0.5, 1.2, hp.S + 1)
    xs[:, 4:] = rs.randn(hp.S + 1, 4)
    us = rs.randn(hp.S, 4) * np.array([1.0, 10.0, 5.0, 20.0])
    Z[:xs.size] = xs.ravel(); Z[xs.size:xs.size + us.size] = us.ravel()
    Z[xs.size + us.size:] = rs.uniform(0, 0.3, hp.M + 2)
    out = dict(Z=Z)
    for method in ('saa', 'baseline'):
        a = HopperOracleA(hp.M, method, 0.2, *feats)
        g = a.g(Z)
        lam = rs.randn(g.size)
        J = a.jac(Z)
        H = a.hess(Z, lam)
        r, c = np.nonzero(J)
        hr, hc = np.nonzero(np.tril(H))
        out.update({method + "_g": g, method + "_lam": lam, method + "_jac_r": r, method + "_jac_c": c,
                    method + "_jac_v": J[r, c], method + "_hess_r": hr, method + "_hess_c": hc,
                    method + "_hess_v": H[hr, hc]})
    np.savez_compressed(os.path.join(HERE, "hopper_M30.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["drone", "car", "hopper"]
    for name in which:
        if name in globals():
            globals()[name]()
            print("minted", name)
