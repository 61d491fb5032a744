"""BASELINE config 1, recorded: the reference's full drone run -- seed 0, first repeat of
sample_uncertain_parameters('saa', M=50), S = 20, alpha in {0.05, 0.1, 0.2, 0.3}
(drone/drone_risk.py:54), 5 warm-up + 60 SCP iterations (:510, :521), with the define / solve split
of drone/drone_times.py:510-542 -- "define" on the GPU path vs on the CPU oracle (Oracle-A, the
reference's algorithm), "solve" on the host QP stand-in (OSQP is not in the image).  One JSON document.

    python examples/config1_drone.py > profiles/config1_drone_r2.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from riskaversetrajopt_b200.drone import drone_params as dp  # noqa: E402
from riskaversetrajopt_b200.drone.drone_risk import Model, L2_error_us  # noqa: E402
from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters  # noqa: E402


def main():
    iters = int(os.environ.get("CONFIG1_ITERS", "60"))
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=dp.M)
    from oracle.oracle_a import DroneOracleA
    out = {"M": dp.M, "S": dp.S, "iterations": iters, "warmup": 5, "solver": "riskaversetrajopt_b200.qp (ADMM, OSQP algorithm)",
           "runs": []}
    for alpha in (0.05, 0.1, 0.2, 0.3):
        model = Model(dp.S, DWs, masses, obs_Qs, 'saa', alpha)
        cpu = DroneOracleA(dp.S, DWs, masses, obs_Qs, 'saa', alpha)
        us_prev = model.initial_guess_us_mat()
        model.define_problem(us_prev)
        for it in range(5):                                  # drone_risk.py:510-517
            model.update_problem(us_prev, it)
            us_prev, _ = model.solve(verbose=False)
        us_prev = model.initial_guess_us_mat()
        define, solve, define_cpu, l2 = [], [], [], []
        for it in range(iters):
            t0 = time.perf_counter()
            model.update_problem(us_prev, it)
            t1 = time.perf_counter()
            us, t_risk = model.solve(verbose=False)
            t2 = time.perf_counter()
            if it % 10 == 0:                                 # the CPU oracle's define on the same iterate
                t3 = time.perf_counter()
                A, l, u = cpu.get_constraints_coeffs(us_prev, it)
                define_cpu.append((time.perf_counter() - t3) * 1e3)
                assert np.allclose(A.data, model.A.data, rtol=1e-8, atol=1e-12)
            define.append((t1 - t0) * 1e3); solve.append((t2 - t1) * 1e3)
            l2.append(float(L2_error_us(us, us_prev)))
            us_prev = us
        sat, Z = model.monte_carlo_constraints(us_prev)
        out["runs"].append({
            "alpha": alpha, "define_ms_median_gpu": float(np.median(define[3:])), "solve_ms_median": float(np.median(solve[3:])),
            "define_ms_median_cpu_oracle_a": float(np.median(define_cpu)),
            "L2_error_trace": [round(x, 6) for x in l2[::5]], "L2_error_final": l2[-1], "t_risk": float(t_risk),
            "training_samples_satisfied": float(np.mean(sat)), "status": model.res.info.status,
            "us_final_first3": np.round(us_prev[:3], 6).tolist()})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
