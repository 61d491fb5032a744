"""Monte-Carlo risk statistics on the device.

The reference validates a solution by rolling out M = 10 000 fresh samples and then
  * V@R_alpha  by sorting the per-sample maxima (drone/drone_main_plot.py:640-652),
  * AV@R_alpha by solving a 2M-row LP in OSQP (drone/drone_risk.py:664-695,
    car/driving.py:639-671, hopper/hopper.py:926-958) and evaluating
    ``t + mean(max(Z - t, 0)) / alpha`` at the LP's ``t``.
Here ``Z`` stays on the GPU (it comes from ``saa_cvar_terms``): V@R is a selection
(``torch.kthvalue``, no sort), and since the minimiser of the Rockafellar-Uryasev objective
``t + E[(Z - t)^+] / alpha`` over a sample is the V@R order statistic itself, AV@R follows
in closed form without any LP.
"""
import math

import torch


def monte_carlo_var(Z, alpha):
    """``sorted(Z)[M - floor(alpha*M) - 1]`` of drone_main_plot.py:640-652, by selection."""
    Z = torch.as_tensor(Z)
    M = Z.numel()
    xth = int(math.floor(alpha * M))
    return float(torch.kthvalue(Z.reshape(-1), M - xth).values.item())   # kthvalue is 1-based


def monte_carlo_avar(Z, alpha, t_risk=None):
    """``t + mean(max(Z - t, 0)) / alpha``; with ``t_risk=None`` the LP's optimal ``t`` (the
    V@R order statistic) is used, which is what the reference's OSQP LP converges to."""
    Z = torch.as_tensor(Z)
    t = monte_carlo_var(Z, alpha) if t_risk is None else float(t_risk)
    return t + float(torch.clamp(Z - t, min=0).mean().item()) / alpha


def fraction_satisfied(Z, tol=1e-6):
    """mean(B_satisfied) with B_satisfied = Z <= tol (drone/drone_risk.py:661, :719)."""
    Z = torch.as_tensor(Z)
    return float((Z <= tol).double().mean().item())
