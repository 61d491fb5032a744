"""Host driver of the CUDA path: owns one ``saa_handle`` and its device buffers.

PyTorch is used for plumbing only (device allocations, streams, pinned host
memory); every computation happens inside ``libsaa_b200.so``.  This is the
layer the per-problem ``Model`` classes (``drone/drone_risk.py``,
``car/driving.py``) sit on; it has no reference counterpart because the
reference does the same work inside JAX (drone/drone_risk.py:282-423).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from ._lib import lib, check

_TORCH_DT = {64: torch.float64, 32: torch.float32}
_NP_DT = {64: np.float64, 32: np.float32}


def _precision_bits(precision):
    if precision in (64, 'fp64', 'float64'):
        return 64
    if precision in (32, 'fp32', 'float32'):
        return 32
    raise ValueError("precision must be 'fp64' or 'fp32'")


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError(
            "riskaversetrajopt_b200 needs a CUDA device (B200, sm_100a); there is no CPU "
            "fallback.  The CPU oracle under oracle/ is test infrastructure only.")
    return torch.device('cuda', torch.cuda.current_device() if device is None else device)


class DevicePath:
    """One problem instance on one GPU.

    ``M_local`` samples starting at ``sample_offset`` of a global set of
    ``M_global``.  ``assemble`` writes into a value array with the CSC layout of
    the matrix for ``M_out`` samples (default: the global matrix).
    """

    def __init__(self, problem, method, S, alpha, M_local, M_global=None, sample_offset=0,
                 variant='risk', precision='fp64', device=None):
        self.device = _require_cuda(device)
        self.bits = _precision_bits(precision)
        self.problem, self.method, self.variant = problem, method, variant
        self.S, self.alpha = int(S), float(alpha)
        self.M_local = int(M_local)
        self.M_global = int(M_local if M_global is None else M_global)
        self.sample_offset = int(sample_offset)
        self._h = C.c_void_p()
        check(lib.saa_create(C.byref(self._h), problem, _lib.METHODS[method], _lib.VARIANTS[variant],
                             self.M_local, self.M_global, self.sample_offset, self.S, self.alpha,
                             self.bits, self.device.index))
        self.M_out, self.first_out = self.M_global, self.sample_offset
        self._patterns = {}
        self._bufs = {}          # relaxed_pattern -> dict(Ax, l, u, const_state)
        self._pinned = {}
        self._host_state = {}    # relaxed_pattern -> (const_state, buffer id) the host copy corresponds to
        self._csc_cache = {}
        self._keep = []          # sample tensors stay alive until the pack kernel ran
        self._params_call = None # (method name, args) of the last set_params_* call (tail.py replays it)
        self.mean_len = int(lib.saa_mean_len(self._h))
        self.mean_sums = torch.zeros(max(self.mean_len, 1), dtype=torch.float64, device=self.device)

    # -- lifetime -----------------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value is not None:
            lib.saa_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, a):
        t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)) if not torch.is_tensor(a) else a
        return t.to(self.device, dtype=torch.float64).contiguous()

    # -- parameters / samples -----------------------------------------------------
    def set_params_drone(self, p, osqp_tol):
        s = _lib.DroneParams()
        s.dt, s.u_max, s.beta = float(p.T / self.S), float(p.u_max), float(p.beta)
        s.drag_coefficient = float(p.drag_coefficient)
        gain = np.asarray(p.feedback_gain, dtype=np.float64)
        gp, gv = -gain[:, :3], -gain[:, 3:]
        if not (np.allclose(gp, gp[0, 0] * np.eye(3)) and np.allclose(gv, gv[0, 0] * np.eye(3))):
            raise ValueError("the CUDA path needs feedback_gain = -[kp*I, kv*I] (axes decouple)")
        s.gain_p, s.gain_v = float(gp[0, 0]), float(gv[0, 0])
        s.x_init[:] = [float(v) for v in p.x_init]
        s.x_final[:] = [float(v) for v in p.x_final]
        s.n_obs = int(p.n_obs)
        for o in range(3):
            for d in range(3):
                s.obs_positions[o][d] = float(p.obs_positions[o][d])
        s.osqp_tol = float(osqp_tol)
        check(lib.saa_set_params_drone(self._h, C.byref(s)), self._h)
        self._params_call = ('set_params_drone', (p, osqp_tol))

    def set_params_car(self, p, beta, osqp_tol):
        s = _lib.CarParams()
        s.dt, s.u_max, s.beta = float(p.dt), float(p.u_max), float(beta)
        s.speed_ped_des = float(p.speed_ped_des)
        s.min_separation_distance = float(p.min_separation_distance)
        s.goal[:] = [float(v) for v in np.concatenate([p.position_ego_goal, p.velocity_ego_goal])]
        s.osqp_tol = float(osqp_tol)
        check(lib.saa_set_params_car(self._h, C.byref(s)), self._h)
        self._params_call = ('set_params_car', (p, beta, osqp_tol))

    def set_samples_drone(self, masses, DWs, obs_Qs):
        Q = obs_Qs if torch.is_tensor(obs_Qs) else np.asarray(obs_Qs)
        if not torch.is_tensor(Q):
            off = Q.copy()
            for d in range(3):
                off[..., d, d] = 0.0
            if np.any(off != 0.0):
                raise ValueError("obs_Qs must be diagonal (as sample_uncertain_parameters builds them)")
        m, dw, q = self._dev(masses), self._dev(DWs), self._dev(obs_Qs)
        assert m.shape == (self.M_local,) and dw.shape == (self.M_local, self.S, 6)
        assert q.shape == (self.M_local, 3, 3, 3)
        check(lib.saa_set_samples_drone(self._h, m.data_ptr(), dw.data_ptr(), q.data_ptr(),
                                        self._stream()), self._h)
        self._keep = [m, dw, q]

    def set_samples_car(self, states_init, omegas_speed, omegas_repulsive, DWs):
        x0, ws, wr, dw = (self._dev(states_init), self._dev(omegas_speed),
                          self._dev(omegas_repulsive), self._dev(DWs))
        assert x0.shape == (self.M_local, 8) and dw.shape == (self.M_local, self.S, 8)
        check(lib.saa_set_samples_car(self._h, x0.data_ptr(), ws.data_ptr(), wr.data_ptr(),
                                      dw.data_ptr(), self._stream()), self._h)
        self._keep = [x0, ws, wr, dw]

    # -- geometry / pattern ---------------------------------------------------------
    def set_output_geometry(self, M_out, first_out):
        check(lib.saa_set_output_geometry(self._h, int(M_out), int(first_out)), self._h)
        self.M_out, self.first_out = int(M_out), int(first_out)
        self._patterns.clear()
        self._bufs.clear()
        self._host_state.clear()
        self._csc_cache.clear()

    def set_active(self, M_active, M_out, first_out):
        """Work on ``M_active`` (<= the capacity the handle was created with) samples, placed at samples
        ``first_out ..`` of a matrix with ``M_out`` samples (``saa_set_active``)."""
        check(lib.saa_set_active(self._h, int(M_active), int(M_out), int(first_out)), self._h)
        if int(M_out) != self.M_out:                       # the pattern depends on M_out only
            self._patterns.clear(); self._bufs.clear(); self._host_state.clear(); self._csc_cache.clear()
        self.M_local, self.M_out, self.first_out = int(M_active), int(M_out), int(first_out)

    def pattern_sizes(self, relaxed_pattern=False):
        r, c, n = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.saa_pattern_sizes(self._h, int(relaxed_pattern), C.byref(r), C.byref(c), C.byref(n)),
              self._h)
        return r.value, c.value, n.value

    def pattern(self, relaxed_pattern=False):
        """-> (n_rows, n_cols, indptr, indices); index dtype int32 like SciPy picks
        while nnz < 2^31, else int64.  Cached (the pattern is static)."""
        key = bool(relaxed_pattern)
        if key not in self._patterns:
            n_rows, n_cols, nnz = self.pattern_sizes(key)
            if max(nnz, n_rows) < 2**31 - 1:
                indptr = np.empty(n_cols + 1, dtype=np.int32)
                indices = np.empty(nnz, dtype=np.int32)
                fn = lib.saa_pattern_i32
            else:
                indptr = np.empty(n_cols + 1, dtype=np.int64)
                indices = np.empty(nnz, dtype=np.int64)
                fn = lib.saa_pattern_i64
            check(fn(self._h, int(key), indptr.ctypes.data, indices.ctypes.data), self._h)
            self._patterns[key] = (n_rows, n_cols, indptr, indices)
        return self._patterns[key]

    # -- buffers ----------------------------------------------------------------------
    def relax_threshold(self):
        return 2 if self.problem == _lib.SAA_DRONE else 1

    def _uses_relaxed_pattern(self, scp_iter):
        return self.problem == _lib.SAA_CAR and scp_iter < 1

    def buffers(self, relaxed_pattern=False):
        key = bool(relaxed_pattern)
        if key not in self._bufs:
            n_rows, _, nnz = self.pattern_sizes(key)
            dt = _TORCH_DT[self.bits]
            self._bufs[key] = dict(
                Ax=torch.empty(nnz, dtype=dt, device=self.device),
                l=torch.empty(n_rows, dtype=dt, device=self.device),
                u=torch.empty(n_rows, dtype=dt, device=self.device),
                const_state=None)
        return self._bufs[key]

    def _us(self, us_mat):
        us = np.ascontiguousarray(np.asarray(us_mat, dtype=np.float64))
        n_u = 3 if self.problem == _lib.SAA_DRONE else 2
        if us.shape != (self.S, n_u):
            raise ValueError(f"us_mat must have shape ({self.S}, {n_u})")
        return us

    # -- the hot path -------------------------------------------------------------------
    def assemble(self, us_mat, scp_iter, Z=None, finalize=True, write_shared=True, out=None):
        """Run linearize+assemble on the current stream.  Returns the dict of device
        buffers {Ax, l, u}.  ``out`` lets the caller supply its own (e.g. peer-mapped)
        buffers as a dict with keys Ax, l, u (tensors or raw pointers) and
        'const_state'."""
        us = self._us(us_mat)
        key = self._uses_relaxed_pattern(scp_iter)
        b = out if out is not None else self.buffers(key)
        ptr = lambda t: t.data_ptr() if torch.is_tensor(t) else int(t)
        relaxed = scp_iter < self.relax_threshold()
        if b.get('const_state') != relaxed:
            check(lib.saa_write_constants(self._h, int(scp_iter), int(write_shared), ptr(b['Ax']),
                                          ptr(b['l']), ptr(b['u']), self._stream()), self._h)
            b['const_state'] = relaxed
        zp = None if Z is None else Z.data_ptr()
        check(lib.saa_linearize_assemble(self._h, us.ctypes.data, int(scp_iter), ptr(b['Ax']),
                                         ptr(b['l']), ptr(b['u']), zp, self.mean_sums.data_ptr(),
                                         int(finalize), self._stream()), self._h)
        return b

    # -- factored record (multi-GPU gather, drone) ------------------------------------------
    def factored_sizes(self):
        n_sp, n_p = C.c_int64(), C.c_int64()
        check(lib.saa_factored_sizes(self._h, C.byref(n_sp), C.byref(n_p)), self._h)
        return n_sp.value, n_p.value

    def linearize_factored(self, us_mat, scp_iter, fsp, fp, u, Z=None):
        """This rank's block as a factored record (sensitivities + trajectory) written to
        ``fsp`` / ``fp`` / ``u`` (tensors or raw, possibly peer-mapped, pointers)."""
        us = self._us(us_mat)
        ptr = lambda t: t.data_ptr() if torch.is_tensor(t) else int(t)
        check(lib.saa_linearize_factored(self._h, us.ctypes.data, int(scp_iter), ptr(fsp), ptr(fp), ptr(u),
                                         None if Z is None else Z.data_ptr(), self.mean_sums.data_ptr(),
                                         self._stream()), self._h)

    def expand_factored(self, scp_iter, fsp, fp, sample_begin, sample_count, Ax):
        ptr = lambda t: t.data_ptr() if torch.is_tensor(t) else int(t)
        check(lib.saa_expand_factored(self._h, int(scp_iter), ptr(fsp), ptr(fp), int(sample_begin),
                                      int(sample_count), ptr(Ax), self._stream()), self._h)

    def write_constants(self, b, scp_iter, write_shared=True):
        """(Re)write the iterate-independent entries if the relaxation state of ``b`` changed."""
        ptr = lambda t: t.data_ptr() if torch.is_tensor(t) else int(t)
        relaxed = scp_iter < self.relax_threshold()
        if b.get('const_state') != relaxed:
            check(lib.saa_write_constants(self._h, int(scp_iter), int(write_shared), ptr(b['Ax']),
                                          ptr(b['l']), ptr(b['u']), self._stream()), self._h)
            b['const_state'] = relaxed

    def finalize_means(self, b, scp_iter=2):
        """After an all-reduce(sum) of ``mean_sums`` over the ranks."""
        ptr = lambda t: t.data_ptr() if torch.is_tensor(t) else int(t)
        check(lib.saa_finalize_means(self._h, self.mean_sums.data_ptr(), int(scp_iter), ptr(b['Ax']),
                                     ptr(b['l']), ptr(b['u']), self._stream()), self._h)

    def _pinned_like(self, name, t):
        p = self._pinned.get(name)
        if p is None or p.shape != t.shape or p.dtype != t.dtype:
            p = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._pinned[name] = p
        return p

    def _iterate_dependent_slices(self, key):
        """Index ranges of (Ax, l, u) that change from one SCP iteration to the next while the
        relaxation state stays the same: the u-column block of A.data (sample-row entries +
        sample-mean entries; the y / slack / t columns are constants), the sample rows' upper bounds
        and the n_fin expectation-row bounds -- 1140 M + 60 M + 123 values for the drone."""
        n_rows, n_cols, indptr, indices = self.pattern(key)
        nu = (3 if self.problem == _lib.SAA_DRONE else 2) * self.S
        n_fin = 6 if self.problem == _lib.SAA_DRONE else 4
        R = (3 if self.problem == _lib.SAA_DRONE else 1) * self.S
        row_s0 = n_fin + (1 + self.M_out if self.method == 'saa' else 0)
        return dict(Ax=[(0, int(indptr[nu]))], l=[(0, n_fin)],
                    u=[(0, n_fin), (row_s0, row_s0 + R * self.M_out)] if not key else [(0, min(8, n_rows))])

    def assemble_host(self, us_mat, scp_iter, copy=True, assembled=None):
        """Assemble and bring (A.data, l, u) to host memory (pinned staging).  The host arrays
        persist between calls: a full device-to-host copy happens only when the relaxation state
        changed (the constant entries were rewritten); otherwise only the iterate-dependent
        slices cross PCIe.  With ``copy=False`` the returned arrays alias the pinned staging
        buffers and are overwritten by the next call.  ``assembled``: the buffer dict of an ``assemble``
        the caller already ran for this iterate (sharded runs finalize the means after an all-reduce)."""
        b = self.assemble(us_mat, scp_iter) if assembled is None else assembled
        key = self._uses_relaxed_pattern(scp_iter)
        outs = [self._pinned_like((name, key), b[name]) for name in ('Ax', 'l', 'u')]
        state = (b['const_state'], id(b['Ax']), outs[0].data_ptr())
        if self._host_state.get(key) != state:
            for h, name in zip(outs, ('Ax', 'l', 'u')):
                h.copy_(b[name], non_blocking=True)
            self._host_state[key] = state
        else:
            sl = self._iterate_dependent_slices(key)
            for h, name in zip(outs, ('Ax', 'l', 'u')):
                for lo, hi in sl[name]:
                    h[lo:hi].copy_(b[name][lo:hi], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        arrs = [h.numpy() for h in outs]
        if copy:
            # private arrays; FP32 storage is widened here (the reference's boundary returns doubles and OSQP
            # wants them) -- a host pass over all values.  copy=False returns the staging buffers as they are
            # (float32 in the FP32 storage mode: half the PCIe bytes and no host pass).
            arrs = [a.astype(np.float64) if self.bits == 32 else a.copy() for a in arrs]
        return arrs

    def d2h_bytes_per_call(self, scp_iter=2):
        """Bytes ``assemble_host`` moves device -> host per call in steady state."""
        key = self._uses_relaxed_pattern(scp_iter)
        sl = self._iterate_dependent_slices(key)
        return sum(hi - lo for v in sl.values() for lo, hi in v) * (self.bits // 8)

    def csc(self, us_mat, scp_iter, copy=True, assembled=None):
        """(A: scipy.sparse.csc_matrix, l, u) -- what ``get_constraints_coeffs`` returns.  With
        ``copy=False`` the matrix object itself is cached per pattern (its ``data`` aliases the
        pinned staging buffer), so a call costs the kernel + the PCIe transfer and nothing else."""
        data, l, u = self.assemble_host(us_mat, scp_iter, copy=copy, assembled=assembled)
        key = self._uses_relaxed_pattern(scp_iter)
        n_rows, n_cols, indptr, indices = self.pattern(key)
        if not copy:
            A = self._csc_cache.get(key)
            if A is None or A.data.ctypes.data != data.ctypes.data:
                A = sp.csc_matrix((data, indices, indptr), shape=(n_rows, n_cols), copy=False)
                if A.data.ctypes.data != data.ctypes.data:  # SciPy made a private copy: re-point it
                    A.data = data
                self._csc_cache[key] = A
            return A, l, u
        A = sp.csc_matrix((data, indices, indptr), shape=(n_rows, n_cols), copy=False)
        return A, l, u

    def check_finite(self, raise_error=True):
        """Samples whose rollout met a zero / non-finite ego-pedestrian distance since the last call
        (the car divides by it, car/driving.py:154; the reference silently returns NaN rows).
        Synchronises the stream.  -> count; raises ``SaaError`` if non-zero and ``raise_error``."""
        n = C.c_int64()
        rc = lib.saa_check_finite(self._h, C.byref(n), self._stream())
        if rc != 0 and (raise_error or n.value == 0):
            check(rc, self._h)
        return n.value

    # -- rollout / CVaR terms ---------------------------------------------------------
    def rollout(self, us_mat):
        us = self._us(us_mat)
        n_x = 6 if self.problem == _lib.SAA_DRONE else 8
        Xs = torch.empty((self.M_local, self.S + 1, n_x), dtype=_TORCH_DT[self.bits], device=self.device)
        check(lib.saa_rollout(self._h, us.ctypes.data, Xs.data_ptr(), self._stream()), self._h)
        return Xs

    def cvar_terms(self, us_mat, t_risk=0.0, sat_tol=1e-6, want_Z=True):
        """-> (Z tensor or None, out3 tensor [sum max(Z-t,0), #{Z<=sat_tol}, max Z]) for
        the local samples (sums; combine across ranks with an all-reduce)."""
        us = self._us(us_mat)
        Z = torch.empty(self.M_local, dtype=_TORCH_DT[self.bits], device=self.device) if want_Z else None
        out3 = torch.empty(3, dtype=torch.float64, device=self.device)
        check(lib.saa_cvar_terms(self._h, us.ctypes.data, float(t_risk), float(sat_tol),
                                 None if Z is None else Z.data_ptr(), out3.data_ptr(),
                                 self._stream()), self._h)
        return Z, out3
