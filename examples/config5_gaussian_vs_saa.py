"""BASELINE config 5: SAA vs Gaussian-approximation constraints, evaluated at M = 10^6 fresh samples.

Controls from (i) the SAA SCP solve (reference defaults: seed 0, M = 50, alpha = 0.1) and (ii) the
deterministic baseline are scored under BOTH models: the Gaussian-linearised chance constraints of
drone/drone_gaussian.py (host: riskaversetrajopt_b200/gaussian.py, uniform risk allocation) and the
sampled CVaR terms of a Monte-Carlo run with 10^6 fresh samples on the GPU.  One JSON document.

    python examples/config5_gaussian_vs_saa.py > profiles/config5_r2.json
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from riskaversetrajopt_b200.drone import drone_params as dp  # noqa: E402
from riskaversetrajopt_b200.drone.drone_risk import Model  # noqa: E402
from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters  # noqa: E402
from riskaversetrajopt_b200.gaussian import score_controls_both_ways  # noqa: E402


def scp(model, iters=30):
    us = model.initial_guess_us_mat()
    model.define_problem(us)
    for it in range(iters):
        model.update_problem(us, it)
        us, t = model.solve(verbose=False)
    return us, t


def main():
    alpha, M_mc = 0.1, int(os.environ.get("CONFIG5_M", "1000000"))
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=dp.M)
    us_saa, t_saa = scp(Model(dp.S, DWs, masses, obs_Qs, 'saa', alpha))
    b = sample_uncertain_parameters('baseline', M=dp.M)
    us_base, _ = scp(Model(dp.S, *b, 'baseline', alpha))
    dev = torch.device("cuda", 0)
    DWm, mm, Qm = bench.synthetic_drone_samples(M_mc, 99, dev)
    mc = Model(dp.S, DWm, mm, Qm, 'saa', alpha)
    out = {"alpha": alpha, "monte_carlo_samples": M_mc, "t_risk_saa": float(t_saa),
           "saa_controls": score_controls_both_ways(mc, us_saa, alpha),
           "baseline_controls": score_controls_both_ways(mc, us_base, alpha)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
