"""One launch of every kernel family at ragged sample counts, for compute-sanitizer:

    compute-sanitizer --tool racecheck python tools/sanitize_run.py
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py

The drone / car kernels hand data between the lanes of a warp through shared memory with
``__syncwarp()`` only (staging -> copy-out, the car's geometry and noise rows); racecheck is the tool
that would see a missing barrier there.  M = 17, 33, 4099: partial tiles, single-sample tails."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riskaversetrajopt_b200.drone import drone_params as dp  # noqa: E402
from riskaversetrajopt_b200.drone.drone_risk import Model as Drone  # noqa: E402
from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters  # noqa: E402
from riskaversetrajopt_b200.car.driving import Model as Car  # noqa: E402
from riskaversetrajopt_b200.hopper import hopper as hp  # noqa: E402

sizes = [int(x) for x in os.environ.get("SAN_M", "17,33,4099").split(",")]
for M in sizes:
    np.random.seed(M)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)
    for prec in ("fp64", "fp32"):
        m = Drone(dp.S, DWs, masses, obs_Qs, 'saa', 0.1, precision=prec)
        us = m.initial_guess_us_mat() + 0.1 * np.random.randn(dp.S, 3)
        for it in (0, 2):
            m.get_constraints_coeffs(us, it)
        m.us_to_state_trajectories(us)
        m.monte_carlo_constraints(us)
        if M >= 17:
            t = m.tail_subproblem(K=max(2, M // 5))
            t.get_constraints_coeffs(us, 2)
        n_sp, n_p = m.path.factored_sizes()
        dt = torch.float64 if prec == "fp64" else torch.float32
        fsp = torch.empty(n_sp, dtype=dt, device="cuda"); fp = torch.empty(n_p, dtype=dt, device="cuda")
        b = m.path.buffers()
        m.path.linearize_factored(us, 2, fsp, fp, b['u'])
        m.path.expand_factored(2, fsp, fp, 0, M, b['Ax'])
        if prec == "fp64":
            # device ADMM: Ruiz sweeps, Gram pass, a few iterations, the termination test -- with the
            # compile-time-horizon kernels, and (SAA_QP_GENERIC=1 in a second run) the run-time-horizon ones
            from riskaversetrajopt_b200.device_qp import DeviceQP
            P, q = m.get_objective_coeffs()
            dq = DeviceQP(m.path, max_iter=20, scaling=2).setup(P, q, m.path.assemble(us, 2))
            dq.solve()
        c = Car(M, 'saa', 0.05, precision=prec)
        usc = c.initial_guess_us_mat() + 0.1 * np.random.randn(20, 2)
        for it in (0, 1, 2):
            c.get_constraints_coeffs(usc, it)
        c.us_to_state_trajectories(usc)
        c.monte_carlo_constraints(usc)
        c.path.check_finite()
        if prec == "fp64":
            P, q = c.get_objective_coeffs()
            dq = DeviceQP(c.path, max_iter=20, scaling=2).setup(P, q, c.path.assemble(usc, 2))
            dq.solve()
        cb = Car(M, 'baseline', 0.05, precision=prec)
        cb.get_constraints_coeffs(usc, 0)
        h = hp.Model(M, 'saa', 0.2, precision=prec)
        Z = np.random.uniform(-1, 1, hp.num_vars(M))
        h.slip_risk_constraints(Z); h.slip_risk_jacobian(Z)
        h.slip_risk_hessian(Z, np.random.randn(h.n_rows)); h.monte_carlo_constraints(Z)
    torch.cuda.synchronize()
    print("ran M =", M, flush=True)
print("sanitize_run: done")
