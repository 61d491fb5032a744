"""Config 2 of BASELINE.json: the car + pedestrian SCP at the reference defaults (seed 0,
M = 50, alpha in {.01,.02,.05,.1} -> here one alpha), timed like car/driving.py:482-513
("define" = linearize + assemble + solver setup/update, "solve" = host QP), the linearization
also timed on the host CPU oracle, then Monte-Carlo validation with fresh samples (:618-703).

    python examples/car_scp.py [--alpha 0.05] [--iters 15]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from riskaversetrajopt_b200.car import driving_params as cp  # noqa: E402
from riskaversetrajopt_b200.car.driving import Model, L2_error_us, sample_uncertain_parameters  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--alpha", type=float, default=0.05)
    ap.add_argument("--iters", type=int, default=15)
    ap.add_argument("--M", type=int, default=cp.M)
    ap.add_argument("--mc", type=int, default=10000)
    args = ap.parse_args()
    np.random.seed(0)                                            # car/driving.py:61
    model = Model(args.M, 'saa', args.alpha)
    us_prev = model.initial_guess_us_mat()
    lin_ms = []
    for scp_iter in range(args.iters):
        t0 = time.perf_counter()
        model.define_problem(us_prev, scp_iter)
        t1 = time.perf_counter()
        us, t_risk = model.solve()
        t2 = time.perf_counter()
        ta = time.perf_counter()
        model.get_constraints_coeffs(us_prev, max(scp_iter, 1), copy=False)
        lin_ms.append((time.perf_counter() - ta) * 1e3)
        print(f"iter {scp_iter:2d}  define {(t1 - t0) * 1e3:7.2f} ms  solve {(t2 - t1) * 1e3:7.2f} ms  "
              f"L2 {L2_error_us(us, us_prev):.3e}  t_risk {t_risk:+.4f}  status {model.res.info.status}")
        us_prev = us
    print(f"GPU linearize+assemble (get_constraints_coeffs, incl. D2H) median {np.median(lin_ms):.3f} ms at M={args.M}")
    try:
        from oracle.oracle_a import CarOracleA
        ref = CarOracleA(model.states_init, model.omegas_speed, model.omegas_repulsive, model.DWs, 'saa', args.alpha)
        ref.get_constraints_coeffs(us_prev, 2)
        t0 = time.perf_counter()
        for _ in range(3):
            ref.get_constraints_coeffs(us_prev, 2)
        print(f"host CPU, reference algorithm (Oracle-A: vmap(jacfwd) + dense packing): "
              f"{(time.perf_counter() - t0) / 3 * 1e3:.1f} ms")
    except Exception as exc:  # the oracle is optional test infrastructure
        print("oracle not available:", exc)
    mc = Model(args.mc, 'saa', args.alpha)
    sat, Z = mc.monte_carlo_constraints(us_prev)
    print(f"Monte-Carlo M={args.mc}: fraction safe {sat.mean():.4f}   AV@R_alpha {mc.monte_carlo_avar(us_prev, t_risk):+.4f}")


if __name__ == "__main__":
    main()
