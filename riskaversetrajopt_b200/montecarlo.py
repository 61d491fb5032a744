"""Monte-Carlo risk statistics on the device.

The reference validates a solution by rolling out M = 10 000 fresh samples and then
  * V@R_alpha  by sorting the per-sample maxima (drone/drone_main_plot.py:640-652),
  * AV@R_alpha by solving a 2M-row LP in OSQP (drone/drone_risk.py:664-695,
    car/driving.py:639-671, hopper/hopper.py:926-958) and evaluating
    ``t + mean(max(Z - t, 0)) / alpha`` at the LP's ``t``.
Here ``Z`` stays on the GPU (it comes from ``saa_cvar_terms`` / ``saa_hopper_cvar_terms``).  V@R is a
selection, done by the library's own exact radix select (``saa_select_tail``: indices of the K
largest values, ascending, ties to the smaller index -- the same kernels that pick the tail
subproblem); since the minimiser of the Rockafellar-Uryasev objective ``t + E[(Z - t)^+] / alpha``
over a sample is that order statistic, AV@R follows in closed form from the selected values
without any LP.  Nothing here calls a library sort / kthvalue.
"""
import ctypes as C
import math

import torch

from ._lib import lib, check


def _largest(path, Z, K):
    """values of the K largest entries of the device tensor ``Z`` (``path``: any DevicePath whose
    ``M_local == Z.numel()`` and storage precision matches; it only lends its scratch)."""
    idx = torch.empty(K, dtype=torch.int64, device=Z.device)
    check(lib.saa_select_tail(path.handle, Z.data_ptr(), int(K), idx.data_ptr(), path._stream()), path.handle)
    return Z[idx]


def monte_carlo_var(path, Z, alpha):
    """``sorted(Z)[M - floor(alpha*M) - 1]`` of drone_main_plot.py:640-652, by selection: the smallest
    of the floor(alpha M) + 1 largest values."""
    M = Z.numel()
    K = int(math.floor(alpha * M)) + 1
    return float(_largest(path, Z.reshape(-1), min(K, M)).min().item())


def monte_carlo_avar(path, Z, alpha, t_risk=None):
    """``t + mean(max(Z - t, 0)) / alpha``; with ``t_risk=None`` the LP's optimal ``t`` (the V@R order
    statistic) is used, which is what the reference's OSQP LP converges to.  Only the selected tail
    contributes to the excess."""
    M = Z.numel()
    K = min(int(math.floor(alpha * M)) + 1, M)
    top = _largest(path, Z.reshape(-1), K)
    t = float(top.min().item()) if t_risk is None else float(t_risk)
    if t_risk is None:
        excess = float((top - t).clamp_min(0).sum().item())          # everything else is <= t
    else:
        excess = float((Z.reshape(-1) - t).clamp_min(0).sum().item())
    return t + excess / (M * alpha)


def fraction_satisfied(Z, tol=1e-6):
    """mean(B_satisfied) with B_satisfied = Z <= tol (drone/drone_risk.py:661, :719)."""
    return float((Z <= tol).double().mean().item())
