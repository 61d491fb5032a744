"""Config 3 of BASELINE.json: the hopper's slip-risk (CVaR no-slip) block at the reference terrain
sampling (seed 1, M = 30, 30 cosine features, 20 contact instants): values, Jacobian and Hessian
contribution on the GPU against the reference's dense autodiff (Oracle-A), and throughput at
M = 10^5 terrain samples.

    python examples/hopper_slip_risk.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from riskaversetrajopt_b200.hopper import hopper as hp  # noqa: E402


def main():
    np.random.seed(1)                                            # hopper/hopper.py:33
    feats = hp.sample_friction_features(hp.M)
    rs = np.random.RandomState(0)
    Z = rs.uniform(-0.3, 0.3, hp.num_vars(hp.M))
    Z[3:(hp.S + 1) * 8:8] = rs.uniform(0.6, 1.1, hp.S + 1)
    model = hp.Model(hp.M, 'saa', 0.2, feats)
    lam = rs.randn(model.n_rows)
    for name, fn in (("g", lambda: model.slip_risk_constraints(Z)), ("jac", lambda: model.slip_risk_jacobian(Z)),
                     ("hess", lambda: model.slip_risk_hessian(Z, lam))):
        fn()
        t0 = time.perf_counter()
        for _ in range(20):
            fn()
        print(f"GPU  {name:5s} M={hp.M}: {(time.perf_counter() - t0) / 20 * 1e3:7.3f} ms per evaluation (incl. host glue)")
    try:
        from oracle.oracle_hopper import HopperOracleA
        a = HopperOracleA(hp.M, 'saa', 0.2, *feats)
        for name, fn in (("g", lambda: a.g(Z)), ("jac", lambda: a.jac(Z)), ("hess", lambda: a.hess(Z, lam))):
            t0 = time.perf_counter()
            fn()
            print(f"CPU  {name:5s} M={hp.M}: {(time.perf_counter() - t0) * 1e3:7.1f} ms (reference algorithm: dense autodiff)")
        assert np.allclose(model.slip_risk_constraints(Z), a.g(Z), rtol=1e-9, atol=1e-12)
        print("values match the oracle")
    except ImportError as exc:
        print("oracle not available:", exc)
    M = 100_000
    big = (0.025 * np.sqrt(2 / 30) * rs.uniform(0, 1, (M, 30)), rs.uniform(0, np.pi, (M, 30)), rs.uniform(0, 2 * np.pi, (M, 30)))
    Zb = rs.uniform(-0.3, 0.3, hp.num_vars(M))
    mb = hp.Model(M, 'saa', 0.2, big)
    mb.slip_risk_constraints(Zb)
    t0 = time.perf_counter()
    mb.slip_risk_constraints(Zb)
    print(f"GPU  g     M={M}: {(time.perf_counter() - t0) * 1e3:7.2f} ms (2 M rows incl. D2H and host packing)")


if __name__ == "__main__":
    main()
