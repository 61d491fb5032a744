"""Config 5 of BASELINE.json (precision part): FP64 vs FP32 of the CUDA path at M = 10^6
synthetic drone samples -- max errors of the Jacobian block, the upper bounds and Z_i, and the
effect on the Monte-Carlo statistics (satisfied fraction, AV@R closed form).

    python examples/precision_study.py [--M 1000000]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from riskaversetrajopt_b200 import _lib  # noqa: E402
from riskaversetrajopt_b200.device_path import DevicePath  # noqa: E402
from riskaversetrajopt_b200.drone import drone_params as dp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=1_000_000)
    ap.add_argument("--alpha", type=float, default=0.1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    M = args.M
    DWs, masses, obs_Qs = bench.synthetic_drone_samples(M, 0, dev)
    us = bench.bench_us()
    out = {}
    res = {}
    for prec in ("fp64", "fp32"):
        p = DevicePath(_lib.SAA_DRONE, 'saa', 20, args.alpha, M, device=0, precision=prec)
        p.set_params_drone(dp, dp.OSQP_TOL)
        p.set_samples_drone(masses, DWs, obs_Qs)
        b = p.assemble(us, 2)
        Z, out3 = p.cvar_terms(us, t_risk=-0.3, sat_tol=1e-6)
        torch.cuda.synchronize()
        n_var = 1140 * M + 177
        res[prec] = dict(Ax=b['Ax'][:n_var].double(), u=b['u'].double(), l6=b['l'][:6].double(),
                         Z=Z.double(), out3=out3.clone())
        del p
    a, f = res["fp64"], res["fp32"]
    scale = float(a["Ax"].abs().max())
    out["M"] = M
    out["max_abs_err_Ax_over_max_abs_Ax"] = float((a["Ax"] - f["Ax"]).abs().max()) / scale
    nz = a["Ax"] != 0
    out["max_rel_err_Ax_per_entry_all_nonzero_entries"] = float(((a["Ax"] - f["Ax"]).abs()[nz] / a["Ax"].abs()[nz]).max())
    out["fp32_mode"] = "FP32 storage of the outputs, FP64 arithmetic (include/saa_b200.h): each entry is the FP64 entry rounded once"
    fin = torch.isfinite(a["u"])
    out["max_abs_err_u"] = float((a["u"][fin] - f["u"][fin]).abs().max())
    out["max_abs_err_mean_rows_l"] = float((a["l6"] - f["l6"]).abs().max())
    out["max_abs_err_Z"] = float((a["Z"] - f["Z"]).abs().max())
    for k, r in res.items():
        o = r["out3"].cpu().numpy()
        out[k] = {"satisfied_fraction": float(o[1]) / M,
                  "avar_closed_form_at_t=-0.3": -0.3 + float(o[0]) / (M * args.alpha),
                  "max_Z": float(o[2])}
    out["satisfied_flips"] = int(((a["Z"] <= 1e-6) != (f["Z"] <= 1e-6)).sum())
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
