"""SCP of the quadrotor CVaR problem with the tail-reduced subproblem: the host
QP solver only sees the K ~ 1.25 alpha M samples with the largest constraint values, selected on
the device at every iterate, so M can be far beyond what the full (68 + 61 M)-row QP allows.
Ends with the Monte-Carlo validation on fresh samples, like examples/drone_scp.py.

The selection is made at the linearisation point, so it is exact only once the iterates have
settled (the "left-out margin" column: <= 0 means no sample outside the selection would have been
active); early iterations with large steps want a generous margin.  The bundled ADMM stand-in for
OSQP is slow on QPs of this size -- keep M modest unless the real osqp package is installed.

    python examples/drone_tail_scp.py [--M 2000] [--alpha 0.1] [--iters 10] [--margin 1.0]
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
from riskaversetrajopt_b200.qp import make_solver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=2000)
    ap.add_argument("--alpha", type=float, default=0.1)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--margin", type=float, default=1.0)
    ap.add_argument("--mc", type=int, default=100000)
    ap.add_argument("--eps", type=float, default=1e-5)
    ap.add_argument("--solver", default=None)
    args = ap.parse_args()
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=args.M)
    model = Model(dp.S, DWs, masses, obs_Qs, 'saa', args.alpha)
    tail = model.tail_subproblem(margin=args.margin)
    print(f"M = {args.M}, K = {tail.K}: QP with {62 + tail.K} variables, {68 + 61 * tail.K} rows "
          f"instead of {62 + args.M} x {68 + 61 * args.M}")
    P, q = tail.get_objective_coeffs(*model.get_objective_coeffs())
    us = model.initial_guess_us_mat()
    A, l, u, idx = tail.get_constraints_coeffs(us, 2)
    solver = make_solver(args.solver)
    solver.setup(P, q, A, l, u, eps_abs=args.eps, eps_rel=args.eps, warm_start=True, polish=False)
    for it in range(args.iters):
        t0 = time.perf_counter()
        A, l, u, idx = tail.get_constraints_coeffs(us, it)
        solver.update(l=l, u=u)
        solver.update(Ax=A.data)
        t1 = time.perf_counter()
        res = solver.solve()
        t2 = time.perf_counter()
        us_new = model.convert_us_vec_to_us_mat(res.x[:60])
        t_risk = res.x[-1]
        err = L2_error_us(us_new, us)
        print(f"iter {it:2d}  define {1e3 * (t1 - t0):7.2f} ms  solve {1e3 * (t2 - t1):8.1f} ms  L2 {err:.3e}  "
              f"t_risk {t_risk:+.4f}  left-out margin {tail.left_out_margin(t_risk):+.3e}  {res.info.status}")
        us = us_new
    # all M training samples at the final controls, then fresh samples
    sat, Z = model.monte_carlo_constraints(us)
    print(f"training samples: fraction safe {sat.mean():.4f}  AV@R_alpha {model.monte_carlo_avar(us, t_risk):+.4f}")
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=args.mc)
    mc = Model(dp.S, DWs, masses, obs_Qs, 'saa', args.alpha)
    sat, Z = mc.monte_carlo_constraints(us)
    print(f"Monte-Carlo M={args.mc}: fraction safe {sat.mean():.4f}   AV@R_alpha {mc.monte_carlo_avar(us, t_risk):+.4f}"
          f"   cost {dp.dt * np.sum(us * us):.4f}")


if __name__ == "__main__":
    main()
