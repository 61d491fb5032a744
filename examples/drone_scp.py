"""Config 1 of BASELINE.json: one full SCP solve of the quadrotor CVaR problem at the
reference defaults (seed 0, M = 50, S = 20), timed like drone/drone_times.py:510-542
("define" = linearize + assemble + solver update, "solve" = host QP), followed by the
Monte-Carlo validation of drone_risk.py:643-725 with M = 10 000 fresh samples.

    python examples/drone_scp.py [--alpha 0.1] [--iters 15] [--solver admm|osqp]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from riskaversetrajopt_b200.drone import drone_params as dp  # noqa: E402
from riskaversetrajopt_b200.drone.drone_risk import Model, L2_error_us  # noqa: E402
from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters  # noqa: E402


def scp(model, iters, solver=None, verbose=True):
    us_prev = model.initial_guess_us_mat()
    model.define_problem(us_prev, solver=solver)
    rows = []
    for scp_iter in range(iters):
        t0 = time.perf_counter()
        model.update_problem(us_prev, scp_iter)
        t1 = time.perf_counter()
        us, t_risk = model.solve(verbose=False)
        t2 = time.perf_counter()
        err = L2_error_us(us, us_prev)
        us_prev = us
        rows.append((scp_iter, (t1 - t0) * 1e3, (t2 - t1) * 1e3, err, t_risk))
        if verbose:
            print(f"iter {scp_iter:2d}  define {rows[-1][1]:7.2f} ms  solve {rows[-1][2]:7.2f} ms  "
                  f"L2 {err:.3e}  t_risk {t_risk:+.4f}")
    return us_prev, t_risk, rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--alpha", type=float, default=0.1)
    ap.add_argument("--iters", type=int, default=15)
    ap.add_argument("--M", type=int, default=dp.M)
    ap.add_argument("--mc", type=int, default=10000)
    ap.add_argument("--solver", default=None)
    args = ap.parse_args()
    np.random.seed(0)                                            # drone_risk.py:57
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=args.M)
    model = Model(dp.S, DWs, masses, obs_Qs, 'saa', args.alpha)
    us, t_risk, rows = scp(model, args.iters, args.solver)
    print("median define ms (iters >= 3):", np.median([r[1] for r in rows[3:]]),
          " median solve ms:", np.median([r[2] for r in rows[3:]]))
    # Monte-Carlo validation with fresh samples (drone_risk.py:646-662, :694)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=args.mc)
    mc = Model(dp.S, DWs, masses, obs_Qs, 'saa', args.alpha)
    sat, Z = mc.monte_carlo_constraints(us)
    print(f"Monte-Carlo M={args.mc}: fraction safe {sat.mean():.4f}   AV@R_alpha {mc.monte_carlo_avar(us, t_risk):+.4f}"
          f"   cost {dp.dt * np.sum(us * us):.4f}")


if __name__ == "__main__":
    main()
