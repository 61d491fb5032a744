"""One-legged hopper on uncertain terrain: the slip-risk (CVaR no-slip) block of the
NLP, backed by the CUDA path.

The reference (``hopper/hopper.py``) solves a direct-transcription NLP with IPOPT
and evaluates ``g``, ``jacrev(g)`` and ``hessian(lambda . g)`` densely with JAX on
every iteration (:566-628).  Only one block of ``g`` depends on the samples:
``slip_risk_constraints`` (:300-367), through the friction field ``mu_i(p_x)``
(:68-81).  This module provides that block – values, Jacobian and Hessian
contribution – with the per-sample work (``n_features`` cosines per sample and
contact instant, and the sample sums of the Hessian) done by
``libsaa_b200.so``; everything else of the reference NLP is sample independent
and stays in the caller's host code (out of scope).

Decision vector layout (reference :106-112):
``Z = [xs (S+1, n_x) row-major | us (S, n_u) row-major | y (M) | slack | t]``.
Row order of the block (saa, :350-366): ``[M alpha t + sum y] [-y_i (M)]
[f_x - mu_i(p) f_z - t - y_i - slack  (i-major, contact-minor)] [0]``;
baseline (:339-348): ``[f_x - mu_i(p) f_z - slack]``.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .._lib import lib, check
from ..device_path import _require_cuda, _precision_bits, _TORCH_DT

# ---- constants of the reference script (hopper/hopper.py:43-69) --------------------
S = 30
M = 30
T = 2.0
dt = T / S
time_jump, time_land = 10, 20
n_x, n_u = 8, 4
mu_nom = 0.10
num_mu_features = 30
CONTACT_STEPS = np.concatenate([np.arange(0, time_jump), np.arange(time_land, S)])   # :305-311


def num_vars(M_):
    return (S + 1) * n_x + S * n_u + M_ + 2


def sample_friction_features(M_, num_features=num_mu_features):
    """Reference draw order on the legacy global RNG (hopper/hopper.py:70-74, seeded with
    ``np.random.seed(1)`` at :33): intensities, thetas, taus, each (M, F)."""
    intensities = np.random.uniform(0, 1, (M_, num_features))
    intensities = np.sqrt(2 / num_features) * intensities
    intensities = 0.025 * intensities
    thetas = np.random.uniform(0, np.pi, (M_, num_features))
    taus = np.random.uniform(0, 2 * np.pi, (M_, num_features))
    return intensities, thetas, taus


class Model:
    """``Model(M, method, alpha)`` as in the reference (:90-104) plus the sampled
    features (the reference reads module globals drawn at import time)."""

    def __init__(self, M_, method='baseline', alpha=0.1, features=None, *, precision='fp64',
                 device=None):
        self.M, self.method, self.alpha = int(M_), method, alpha
        if features is None:
            features = sample_friction_features(self.M)
        I, th, ta = (np.ascontiguousarray(f, dtype=np.float64) for f in features)
        if method == 'baseline':                      # :96-99
            I, th, ta = 0 * I, 0 * th, 0 * ta
        self.intensities, self.thetas, self.taus = I, th, ta
        self.device = _require_cuda(device)
        self.bits = _precision_bits(precision)
        self._h = C.c_void_p()
        check(lib.saa_create(C.byref(self._h), _lib.SAA_HOPPER, _lib.METHODS[method], 0, self.M,
                             self.M, 0, S, float(alpha), self.bits, self.device.index))
        d = lambda a: torch.as_tensor(a).to(self.device)
        self._feat = [d(I), d(th), d(ta)]
        check(lib.saa_set_samples_hopper(self._h, I.shape[1], mu_nom, self._feat[0].data_ptr(),
                                         self._feat[1].data_ptr(), self._feat[2].data_ptr(),
                                         self._stream()), self._h)
        self.n_c = len(CONTACT_STEPS)
        tdt = _TORCH_DT[self.bits]
        self._mu = torch.empty(self.M * self.n_c, dtype=tdt, device=self.device)
        self._dmu = torch.empty(self.M * self.n_c, dtype=tdt, device=self.device)
        self._hs = torch.empty(2 * self.n_c, dtype=torch.float64, device=self.device)
        self._build_structure()

    def __del__(self):
        try:
            if self._h.value is not None:
                lib.saa_destroy(self._h)
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- layout helpers ------------------------------------------------------------
    @property
    def n_rows(self):
        nc = self.n_c
        return self.M * nc if self.method == 'baseline' else 1 + self.M + self.M * nc + 1

    def _idx(self):
        M_ = self.M
        ix = lambda t, k: t * n_x + k
        iu = lambda t, k: (S + 1) * n_x + t * n_u + k
        iy0 = (S + 1) * n_x + S * n_u
        return ix, iu, iy0, iy0 + M_, iy0 + M_ + 1

    def _contact_geometry(self, Z):
        """end-effector x positions (:166-171), forces and dp/d(x0,x2,x3) at the contact instants."""
        Z = np.asarray(Z, dtype=np.float64)
        xs = Z[:(S + 1) * n_x].reshape(S + 1, n_x)
        us = Z[(S + 1) * n_x:(S + 1) * n_x + S * n_u].reshape(S, n_u)
        t = CONTACT_STEPS
        x0, x2, x3 = xs[t, 0], xs[t, 2], xs[t, 3]
        px = x0 + x3 * np.sin(x2)
        return px, us[t, 2], us[t, 3], x2, x3

    def _point(self, Z):
        """``saa_hopper_point`` of the decision vector: the O(n_c) contact geometry (host)."""
        Z = np.asarray(Z, dtype=np.float64)
        px, fx, fz, x2, x3 = self._contact_geometry(Z)
        _, _, _, islack, it = self._idx()
        pt = _lib.HopperPoint()
        pt.n_c = self.n_c
        for name, arr in (("px", px), ("fx", fx), ("fz", fz), ("x2", x2), ("x3", x3)):
            getattr(pt, name)[:self.n_c] = [float(v) for v in arr]
        pt.t_risk, pt.slack = float(Z[it]), float(Z[islack])
        return pt

    def _friction(self, px, lam=None):
        """Device evaluation: mu, mu' (M, n_c) and, with multipliers, the per-contact sums
        sum_i lam_ic mu_i', sum_i lam_ic mu_i''."""
        px = np.ascontiguousarray(px, dtype=np.float64)
        lam_t = None
        if lam is not None:
            lam_t = torch.as_tensor(np.ascontiguousarray(lam, dtype=np.float64)).to(self.device)
        check(lib.saa_hopper_friction(self._h, self.n_c, px.ctypes.data, self._mu.data_ptr(),
                                      self._dmu.data_ptr(),
                                      None if lam_t is None else lam_t.data_ptr(),
                                      None if lam_t is None else self._hs.data_ptr(),
                                      self._stream()), self._h)
        mu = self._mu.cpu().numpy().astype(np.float64).reshape(self.M, self.n_c)
        dmu = self._dmu.cpu().numpy().astype(np.float64).reshape(self.M, self.n_c)
        hs = None if lam is None else self._hs.cpu().numpy().reshape(self.n_c, 2)
        return mu, dmu, hs

    # ---- values: reference slip_risk_constraints (:300-367), assembled on the device ----
    def slip_risk_constraints_device(self, Z):
        """-> device tensor (n_rows,) in the handle's storage precision."""
        Z = np.asarray(Z, dtype=np.float64)
        pt = self._point(Z)
        _, _, iy0, _, _ = self._idx()
        y = None
        if self.method == 'saa':
            y = torch.as_tensor(np.ascontiguousarray(Z[iy0:iy0 + self.M])).to(self.device)
        g = torch.empty(self.n_rows, dtype=_TORCH_DT[self.bits], device=self.device)
        check(lib.saa_hopper_g(self._h, C.byref(pt), None if y is None else y.data_ptr(), g.data_ptr(),
                               self._stream()), self._h)
        return g

    def slip_risk_constraints(self, Z):
        return self.slip_risk_constraints_device(Z).cpu().numpy().astype(np.float64)

    # ---- Jacobian: the rows jacrev(g) holds for this block (:568-569) -------------------
    def _build_structure(self):
        M_, nc = self.M, self.n_c
        ix, iu, iy0, islack, it = self._idx()
        t = CONTACT_STEPS
        i = np.arange(M_)
        r0 = 0 if self.method == 'baseline' else 1 + M_
        rows_ic = r0 + i[:, None] * nc + np.arange(nc)[None, :]          # (M, n_c)
        cols5 = np.stack([ix(t, 0), ix(t, 2), ix(t, 3), iu(t, 2), iu(t, 3)], axis=0)   # (5, n_c)
        rows, cols = [np.broadcast_to(rows_ic[None], (5, M_, nc)).ravel()], \
                     [np.broadcast_to(cols5[:, None, :], (5, M_, nc)).ravel()]
        rows.append(rows_ic.ravel()); cols.append(np.full(M_ * nc, islack))
        self._n_var_entries = 5 * M_ * nc
        if self.method == 'saa':
            rows.append(rows_ic.ravel()); cols.append(np.broadcast_to((iy0 + i)[:, None], (M_, nc)).ravel())
            rows.append(rows_ic.ravel()); cols.append(np.full(M_ * nc, it))
            rows.append(np.zeros(M_, dtype=np.int64)); cols.append(iy0 + i)       # row 0: d/dy_i = 1
            rows.append(np.zeros(1, dtype=np.int64)); cols.append(np.array([it]))  # row 0: d/dt = M alpha
            rows.append(1 + i); cols.append(iy0 + i)                              # rows 1+i: -1
        self.jac_rows = np.concatenate(rows).astype(np.int64)
        self.jac_cols = np.concatenate(cols).astype(np.int64)
        # constant part of the values (everything but the four iterate-dependent arrays)
        N = M_ * nc
        consts = [np.ones(N), None, -np.ones(N)]            # d/df_x, (d/df_z: varying), d/dslack
        if self.method == 'saa':
            consts += [-np.ones(N), -np.ones(N), np.ones(M_), np.array([M_ * self.alpha]), -np.ones(M_)]
        self._jac_consts = consts
        self._jac_dev = torch.empty(4 * N, dtype=_TORCH_DT[self.bits], device=self.device)
        self._hess_dev = torch.empty(10 * nc, dtype=torch.float64, device=self.device)

    def slip_risk_jacobian_device(self, Z):
        """-> device tensor [4][M n_c]: d row / d(x0, x2, x3, f_z) of the sample rows (the entries
        that depend on the iterate; ``saa_hopper_jac``)."""
        pt = self._point(Z)
        check(lib.saa_hopper_jac(self._h, C.byref(pt), self._jac_dev.data_ptr(), self._stream()), self._h)
        return self._jac_dev

    def slip_risk_jacobian(self, Z):
        """-> (rows, cols, vals): COO triplets of d(slip_risk_constraints)/dZ.  rows/cols are
        static (``self.jac_rows``, ``self.jac_cols``)."""
        N = self.M * self.n_c
        v = self.slip_risk_jacobian_device(Z).cpu().numpy().astype(np.float64).reshape(4, N)
        c = self._jac_consts
        vals = [v[0], v[1], v[2], c[0], v[3]] + c[2:]
        return self.jac_rows, self.jac_cols, np.concatenate(vals)

    # ---- Hessian of lambda . g restricted to this block (:571-575) ----------------------
    def slip_risk_hessian(self, Z, lagrange):
        """``lagrange``: multipliers of this block's rows (length n_rows).
        -> (rows, cols, vals) of the symmetric Hessian contribution, lower triangle
        (rows >= cols), 10 entries per contact instant, formed on the device (``saa_hopper_hess``)."""
        M_, nc = self.M, self.n_c
        lagrange = np.asarray(lagrange, dtype=np.float64)
        r0 = 0 if self.method == 'baseline' else 1 + M_
        lam = torch.as_tensor(np.ascontiguousarray(lagrange[r0:r0 + M_ * nc])).to(self.device)
        pt = self._point(Z)
        check(lib.saa_hopper_hess(self._h, C.byref(pt), lam.data_ptr(), self._hess_dev.data_ptr(),
                                  self._stream()), self._h)
        vals = self._hess_dev.cpu().numpy()
        ix, iu, *_ = self._idx()
        t = CONTACT_STEPS
        var = np.stack([ix(t, 0), ix(t, 2), ix(t, 3), iu(t, 3)], axis=0)       # (4, n_c): x0, x2, x3, f_z
        rows, cols = [], []
        for a in range(4):
            for b in range(a + 1):
                rows.append(var[a]); cols.append(var[b])
        return np.concatenate(rows), np.concatenate(cols), vals

    # ---- Monte-Carlo verification (:910-925, :957) ------------------------------------------
    def monte_carlo_constraints(self, Z, t_risk=None, sat_tol=1e-6):
        """-> (B_satisfied (M,), max_constraint (M,), out3) of ``no_slip_constraints_verification``
        vmapped over this model's samples at the trajectory in ``Z``; out3 = [sum max(Z_i - t, 0),
        #{Z_i <= sat_tol}, max Z_i] with t = ``t_risk`` (default: Z's t)."""
        pt = self._point(Z)
        if t_risk is not None:
            pt.t_risk = float(t_risk)
        Zi = torch.empty(self.M, dtype=_TORCH_DT[self.bits], device=self.device)
        out3 = torch.empty(3, dtype=torch.float64, device=self.device)
        check(lib.saa_hopper_cvar_terms(self._h, C.byref(pt), float(sat_tol), Zi.data_ptr(), out3.data_ptr(),
                                        self._stream()), self._h)
        Zh = Zi.cpu().numpy().astype(np.float64)
        return Zh <= sat_tol, Zh, out3.cpu().numpy()

    def monte_carlo_avar(self, Z, t_risk, alpha=None):
        """t + mean(max(Z_i - t, 0)) / alpha (closed form at :957)."""
        alpha = self.alpha if alpha is None else alpha
        _, _, out3 = self.monte_carlo_constraints(Z, t_risk)
        return t_risk + float(out3[0]) / (self.M * alpha)

    # ---- IPOPT-callback style helpers (write into caller-owned arrays, :586-628) ---------
    def eval_g_into(self, x, out, row_offset):
        out[row_offset:row_offset + self.n_rows] = self.slip_risk_constraints(x)
        return out

    def eval_jac_g_into(self, x, out_dense, row_offset):
        """``out_dense``: the reference's dense (ncon, nvar) Jacobian (or a view of it)."""
        r, c, v = self.slip_risk_jacobian(x)
        out_dense[row_offset:row_offset + self.n_rows, :] = 0.0
        np.add.at(out_dense, (r + row_offset, c), v)
        return out_dense

    def eval_h_add(self, x, lagrange_block, out_dense):
        """Add this block's contribution to a dense symmetric (nvar, nvar) Hessian."""
        r, c, v = self.slip_risk_hessian(x, lagrange_block)
        np.add.at(out_dense, (r, c), v)
        off = r != c
        np.add.at(out_dense, (c[off], r[off]), v[off])
        return out_dense
