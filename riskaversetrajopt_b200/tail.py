"""Tail-reduced CVaR subproblem (no counterpart in the reference).

The Rockafellar-Uryasev program of the reference (drone/drone_risk.py:327-368,
car/driving.py:331-372) has one auxiliary variable ``y_i`` and ``S*n_obs`` rows per
sample.  At its solution ``y_i = max(g_i - t, 0)`` is non-zero only on the upper
alpha-tail of ``Z_i = max g_i``, so the host QP solver can be handed the
``K ~ (1 + margin) alpha M`` samples with the largest ``Z_i`` at the current iterate:
the matrix is the full one with the other samples' rows and ``y`` columns deleted
(the CVaR row keeps ``M alpha t``, the expectation rows keep the mean over ALL ``M``
samples).  Everything sample-sized runs on the device:

  means + Z_i of all M samples   saa_linearize_means   (no matrix is written)
  K largest Z_i, ascending       saa_select_tail       (radix select, deterministic)
  packed inputs of those         saa_gather_samples
  linearize + assemble K         saa_linearize_assemble on a (K of M) handle

The reduction is exact whenever the samples left out satisfy ``g_i <= t`` at the
subproblem's solution; ``TailSubproblem.left_out_margin`` reports that a posteriori.
"""
import ctypes as C
import math

import numpy as np
import scipy.sparse as sp
import torch

from ._lib import lib, check
from .device_path import DevicePath, _TORCH_DT


class TailSubproblem:
    def __init__(self, path, K=None, margin=0.25, out_geometry=None):
        """``path``: the ``DevicePath`` of a drone / car ``Model`` (method 'saa').  ``K`` of its
        local samples are kept (default ``ceil((1 + margin) * alpha * M_local)``).
        ``out_geometry = (K_total, first)`` places them as samples ``first .. first + K`` of a
        matrix for ``K_total`` samples (multi-GPU: ``dist.ShardedTailAssembler``); by default the
        matrix holds exactly these K samples, which requires all samples to be on this GPU."""
        if path.method != 'saa':
            raise ValueError("the tail reduction applies to the CVaR ('saa') program")
        if out_geometry is None and path.M_local != path.M_global:
            raise ValueError("the samples are sharded: use dist.ShardedTailAssembler")
        if path._params_call is None:
            raise ValueError("set the parameters and samples of the full path first")
        self.full = path
        M = path.M_local
        self.K = int(min(M, max(1, math.ceil((1.0 + margin) * path.alpha * M))) if K is None else K)
        if not 1 <= self.K <= M:
            raise ValueError("need 1 <= K <= M")
        self.sub = DevicePath(path.problem, path.method, path.S, path.alpha, self.K, M_global=path.M_global,
                              sample_offset=0, variant=path.variant,
                              precision={64: 'fp64', 32: 'fp32'}[path.bits], device=path.device.index)
        self.sub.set_output_geometry(*(out_geometry if out_geometry is not None else (self.K, 0)))
        name, args = path._params_call
        getattr(self.sub, name)(*args)
        dev = path.device
        self.Z = torch.empty(M, dtype=_TORCH_DT[path.bits], device=dev)
        self.idx = torch.empty(self.K, dtype=torch.int64, device=dev)

    # -- device side ------------------------------------------------------------------
    def select(self, us_mat):
        """Means + Z_i of all local samples (``full.mean_sums``, ``self.Z``), the K largest
        (``self.idx``, ascending) and their packed inputs into the K-sample handle."""
        f, s = self.full, self.sub
        us = f._us(us_mat)
        st = f._stream()
        check(lib.saa_linearize_means(f._h, us.ctypes.data, self.Z.data_ptr(), f.mean_sums.data_ptr(), st), f._h)
        check(lib.saa_select_tail(f._h, self.Z.data_ptr(), self.K, self.idx.data_ptr(), st), f._h)
        check(lib.saa_gather_samples(s._h, f._h, self.idx.data_ptr(), st), s._h)
        return us

    def assemble(self, us_mat, scp_iter, out=None, write_shared=True, finalize=True):
        """-> dict of device buffers {Ax, l, u} of the K-sample matrix; ``self.idx`` holds the
        selected sample indices (ascending), ``self.Z`` the Z_i of all samples.  ``out`` /
        ``write_shared`` as in ``DevicePath.assemble``; with ``finalize=False`` the expectation
        rows are left to the caller (``finalize_means`` after the all-reduce of ``full.mean_sums``)."""
        us = self.select(us_mat)
        b = self.sub.assemble(us, scp_iter, finalize=False, write_shared=write_shared, out=out)
        if finalize:
            self.finalize_means(b, scp_iter)
        return b

    def finalize_means(self, b, scp_iter):
        """Expectation rows = mean over ALL samples: the sums of the means pass over the full sample
        set (all-reduced across ranks by the caller if sharded), not those of the K-sample launch."""
        f, s = self.full, self.sub
        ptr = lambda t: t.data_ptr() if torch.is_tensor(t) else int(t)
        check(lib.saa_finalize_means(s._h, f.mean_sums.data_ptr(), int(scp_iter), ptr(b['Ax']),
                                     ptr(b['l']), ptr(b['u']), f._stream()), s._h)

    # -- host side ----------------------------------------------------------------------
    def get_constraints_coeffs(self, us_mat, scp_iter, copy=True):
        """-> (A csc, l, u, idx): the reduced counterpart of ``Model.get_constraints_coeffs``.
        QP variables: (u, y[idx], slack, t)."""
        s = self.sub
        b = self.assemble(us_mat, scp_iter)
        # the constant entries (y / slack / t columns, constant bounds) do not depend on WHICH samples
        # were selected, so after the first call only the iterate-dependent slices cross PCIe
        A, l, u = s.csc(us_mat, scp_iter, copy=copy, assembled=b)
        idx = self.idx.cpu().numpy()
        return A, l, u, idx

    def d2h_bytes_per_call(self, scp_iter=2):
        return self.sub.d2h_bytes_per_call(scp_iter) + 8 * self.K

    def get_objective_coeffs(self, P_full, q_full):
        """Restrict the full problem's (P, q) to the variables (u, y[idx], slack, t).  The y block of
        P and q is zero (drone/drone_risk.py:376-391), so this does not depend on ``idx``."""
        nu = P_full.shape[0] - self.full.M_local - 2
        keep = np.concatenate([np.arange(nu), nu + np.arange(self.K), [P_full.shape[0] - 2, P_full.shape[0] - 1]])
        P = sp.csc_matrix(P_full)[keep][:, keep]
        return sp.csc_matrix(P), np.asarray(q_full)[keep]

    def left_out_margin(self, t_risk):
        """max over the samples NOT selected of (Z_i - t) at the iterate of the last ``assemble``:
        <= 0 means the reduction was exact there (their y_i would be 0)."""
        mask = torch.ones_like(self.Z, dtype=torch.bool)
        mask[self.idx[:self.K]] = False
        if not bool(mask.any()):
            return -math.inf
        return float((self.Z[mask].max() - t_risk).item())
