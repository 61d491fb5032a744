"""Device-resident ADMM for the CVaR program of one SCP iteration (second option).

The reference hands ``(P, q, A, l, u)`` to OSQP on the host (drone/drone_risk.py:425-461,
car/driving.py:423-447).  At M >= 10^5 samples that hand-off is the bottleneck: 9.6 KB per sample cross
PCIe and no host solver ingests a 61 M-row matrix.  ``DeviceQP`` solves the same QP where the matrix
already is -- the ``Ax / l / u`` buffers ``DevicePath.assemble`` wrote -- with OSQP's algorithm (the one
``qp.OSQPLike`` implements on the host: ADMM, Ruiz equilibration, per-row rho, over-relaxation alpha,
adaptive rho, OSQP's termination test) and a structure-exploiting linear solve: see
``csrc/qp_kernels.cuh``.  Per ADMM iteration: one pass over the samples (one warp per sample; the
Jacobian values are read in place), a reduction of nu + 4 numbers (all-reduced across ranks when the
samples are sharded -- this is the consumer for which the row blocks can stay sharded) and a one-block
kernel for the nu + 2 dense variables.  The host keeps the O(nu^2) logic: inverting the Schur
complement when rho changes, the Ruiz updates of the dense columns / global rows, the termination test
on reduced numbers.  Iterates agree with ``qp.OSQPLike`` to rounding (tests/test_gpu_qp.py).

Only ``u``, ``slack``, ``t`` (and ``y`` on request) ever leave the device.
"""
import ctypes as C
import time
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

_INF = 1e20
_SCALE, _GRAM, _PASS, _CHECK = 0, 1, 2, 3


def _guard(v):
    v = np.asarray(v, dtype=np.float64)
    return np.minimum(np.where(v < 1e-4, 1.0, v), 1e4)


class _Layout:
    """The flat int64 array of ``saa_qp_layout`` with names."""
    _HEAD = ("n_fin", "nu", "R", "S", "blk", "M_out", "first_out", "row_cvar", "row_y0", "row_s0", "row_slack",
             "row_ctrl0", "ycol0", "slackcol", "tcol", "nnz", "n_rows", "nact", "nnzJ", "npairs")
    _DENSE = ("Fs", "ctS", "slS", "vcw", "rho", "lo", "hi", "z", "lam", "Sinv", "h", "pw", "scal", "qw", "xw", "xt",
              "total")

    def __init__(self, path):
        buf = np.zeros(1 << 16, dtype=np.int64)
        check(lib.saa_qp_layout(path.handle, buf.ctypes.data, buf.size), path.handle)
        it = iter(buf.tolist())
        for k in self._HEAD:
            setattr(self, k, next(it))
        self.off = {k: next(it) for k in self._DENSE}
        self.ucol, self.fin = [], []           # element offset of each u column; its (final row, offset) pairs
        self.run_len = []
        for c in range(self.nu):
            uc, nf = next(it), next(it)
            rows = [next(it) for _ in range(4)]
            self.ucol.append(uc)
            self.fin.append([(rows[k], uc + k) for k in range(nf)])
            self.run_len.append(next(it))
        self.active = [next(it) for _ in range(self.nact)]
        self.pairs = [(next(it), next(it)) for _ in range(self.npairs)]
        self.nw = self.nu + 2
        self.ng = self.n_fin + 2 + self.nu


class DeviceQP:
    def __init__(self, path, group=None, sharded=None, eps_abs=1e-3, eps_rel=1e-3, max_iter=20000, rho=0.1, sigma=1e-6, alpha=1.6,
                 scaling=10, adaptive_rho_interval=50, adaptive_rho_tolerance=5.0, check_interval=10, polish=False,
                 verbose=False):
        if path.method != 'saa' or path.bits != 64:
            raise ValueError("DeviceQP solves the FP64 CVaR ('saa') program")
        if polish:                 # as qp.OSQPLike: polishing is emulated by iterating further
            eps_abs, eps_rel, max_iter = min(eps_abs, 1e-6), min(eps_rel, 1e-6), max(max_iter, 200000)
        # sharded: every rank holds a compact block of its own samples (M_out = its count) and the QP is over
        # the union; otherwise the QP is over the M_out samples of this GPU's matrix (the full set, or the K
        # samples of a tail-reduced subproblem)
        self.path, self.group = path, group
        self._sharded = bool(group is not None) if sharded is None else bool(sharded)
        self.o = SimpleNamespace(eps_abs=eps_abs, eps_rel=eps_rel, max_iter=max_iter, sigma=sigma, alpha=alpha,
                                 scaling=scaling, adaptive_rho_interval=adaptive_rho_interval,
                                 adaptive_rho_tolerance=float(adaptive_rho_tolerance),
                                 check_interval=check_interval, verbose=verbose)
        self.rho = rho
        self.dev = path.device
        self.launches = 0
        self._b = None

    # -- small helpers ------------------------------------------------------------------------
    def _dev(self, a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(self.dev)

    def _all_reduce(self, t, n_max):
        if not self._sharded:
            return
        import torch.distributed as dist
        if n_max > 0:
            dist.all_reduce(t[:n_max], op=dist.ReduceOp.MAX, group=self.group)
        if n_max < t.numel():
            dist.all_reduce(t[n_max:], op=dist.ReduceOp.SUM, group=self.group)

    def _reduced(self, kind, n_max):
        """Reduce the partials of the last pass of ``kind`` (+ across ranks) -> device vector."""
        part, nblk, plen = self._part[kind]
        out = self._red[kind]
        check(lib.saa_qp_reduce(self.path.handle, part.data_ptr(), nblk, plen, n_max, out.data_ptr(),
                                self.path._stream()), self.path.handle)
        self.launches += 1
        self._all_reduce(out, n_max if n_max < plen else plen)
        return out

    # -- setup ----------------------------------------------------------------------------------
    def setup(self, P, q, b):
        """``P, q``: the objective (``Model.get_objective_coeffs``); ``b``: the assembled device buffers
        (``DevicePath.assemble``: dict with Ax, l, u).  Computes the Ruiz scaling (kept for later
        ``update`` calls, as OSQP does) and the first factorisation."""
        p = self.path
        L = self.L = _Layout(p)
        nu, nw, M, R = L.nu, L.nw, p.M_local, L.R
        if self._sharded:
            import torch.distributed as dist
            if self.group is None:
                self.group = dist.group.WORLD
            cnt = torch.tensor([M], dtype=torch.int64, device=self.dev)
            dist.all_reduce(cnt, group=self.group)
            self.M_qp = int(cnt.item())
        else:
            if M != L.M_out:
                raise ValueError("this GPU's matrix holds other ranks' samples too: pass group= / sharded=True "
                                 "with compact per-rank blocks (set_output_geometry(M_local, 0))")
            self.M_qp = M
        n = self.n = nu + self.M_qp + 2
        if P.shape[0] != n:
            raise ValueError(f"P must be the objective over (u, y[{self.M_qp}], slack, t)")
        # only the blocks the reference's objective has (drone_risk.py:376-391): u block, slack and t diagonal;
        # never densify P (n ~ 10^5..10^6)
        blk = lambda a: np.asarray(a.todense() if hasattr(a, 'todense') else a, dtype=np.float64)
        self.P_uu = blk(P[:nu, :nu])
        self.P_ss, self.P_tt = float(P[n - 2, n - 2]), float(P[n - 1, n - 1])
        if hasattr(P, 'nnz') and P.nnz != np.count_nonzero(self.P_uu) + (self.P_ss != 0) + (self.P_tt != 0):
            raise ValueError("DeviceQP expects P = blkdiag(P_uu, 0_y, p_slack, p_t)")
        q = np.asarray(q, dtype=np.float64)
        self.q_w = np.concatenate([q[:nu], [q[n - 2], q[n - 1]]])
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.st = {k: torch.zeros(M * (R if k in ('Es', 'zs', 'ls') else 1), **f64)
                   for k in ("Dy", "Ey", "Es", "xy", "rloc", "zy", "ly", "zs", "ls")}
        for k in ("Dy", "Ey", "Es"):
            self.st[k].fill_(1.0)
        self._st = _lib.QpSampleState(**{k: v.data_ptr() for k, v in self.st.items()})
        self._part, self._red = {}, {}
        for kind in (_SCALE, _GRAM, _PASS, _CHECK):
            nb, pl = C.c_int64(), C.c_int64()
            check(lib.saa_qp_partials(p.handle, kind, C.byref(nb), C.byref(pl)), p.handle)
            self._part[kind] = (torch.zeros(nb.value * pl.value, **f64), nb.value, pl.value)
            self._red[kind] = torch.zeros(pl.value, **f64)
        self.G = torch.zeros(L.off['total'], **f64)
        # gather indices of the sample-independent entries
        fin_idx = [o for c in range(nu) for (_, o) in L.fin[c]]
        self._fin_rc = [(r, c) for c in range(nu) for (r, _) in L.fin[c]]
        ctrl_idx = [(L.ucol[c + 1] if c + 1 < nu else L.ycol0) - 1 for c in range(nu)]
        g0 = L.first_out
        const_idx = [L.slackcol, L.tcol, L.slackcol + 1 + L.M_out,                       # cvar_s, cvar_t, sl
                     L.ycol0 + g0 * (2 + R), L.ycol0 + g0 * (2 + R) + 1, L.ycol0 + g0 * (2 + R) + 2,   # cvar_y, yd, yr
                     L.slackcol + 1 + g0, L.tcol + 1 + g0 * R]                           # ys, tr
        self._ax_idx = torch.as_tensor(fin_idx + ctrl_idx + const_idx, dtype=torch.int64, device=self.dev)
        rows = list(range(L.n_fin)) + [L.row_cvar, L.row_slack] + [L.row_ctrl0 + c for c in range(nu)]
        self._row_idx = torch.as_tensor(rows, dtype=torch.int64, device=self.dev)
        self._b = b
        self._read_globals()
        self._ruiz()
        self._factor()
        self._stale = True
        return self

    def _read_globals(self):
        """The sample-independent pieces of (A, l, u): final rows, control rows, constants, bounds."""
        L, b = self.L, self._b
        nu = L.nu
        if self.path.M_local == 0:
            raise ValueError("a rank without samples cannot read the matrix constants")
        v = b['Ax'][self._ax_idx].cpu().numpy()
        nfin = len(self._fin_rc)
        self.F = np.zeros((L.n_fin, nu))
        for (r, c), x in zip(self._fin_rc, v[:nfin]):
            self.F[r, c] = x
        self.ctrl = v[nfin:nfin + nu].copy()
        (self.cvar_s, self.cvar_t, self.sl, self.cvar_y, self.yd, self.yr, self.ys, self.tr) = v[nfin + nu:]
        lo, hi = b['l'][self._row_idx].cpu().numpy(), b['u'][self._row_idx].cpu().numpy()
        self.lo_g = np.maximum(np.where(np.isnan(lo), -_INF, lo), -_INF)
        self.hi_g = np.minimum(np.where(np.isnan(hi), _INF, hi), _INF)

    def _launch(self, kind, *extra):
        p, b = self.path, self._b
        part = self._part[kind][0]
        ptr = lambda t: t.data_ptr() if torch.is_tensor(t) else int(t)
        head = (p.handle, ptr(b['Ax']), ptr(b['l']), ptr(b['u']), self._Dw_dev.data_ptr())
        st, stream = C.byref(self._st), p._stream()
        if kind == _SCALE:
            check(lib.saa_qp_scale_pass(*head, self.Ec, st, part.data_ptr(), stream), p.handle)
        elif kind == _GRAM:
            check(lib.saa_qp_gram_pass(*head, self.Ec, self.rho, self.o.sigma, st, part.data_ptr(), stream), p.handle)
        elif kind == _PASS:
            xt = self.G[self.L.off['xt']:]
            check(lib.saa_qp_admm_pass(*head, self.Ec, self.rho, self.o.sigma, self.o.alpha, st, xt.data_ptr(),
                                       int(extra[0]), part.data_ptr(), stream), p.handle)
        else:
            check(lib.saa_qp_check_pass(*head, self.Ec, st, extra[0].data_ptr(), part.data_ptr(), stream), p.handle)
        self.launches += 1

    # -- Ruiz equilibration (qp.py:_scale; oracle/arrow_admm.py:_scale) -----------------------------
    def _ruiz(self):
        L = self.L
        nu, n = L.nu, self.n
        self.Du, self.Ds, self.Dt = np.ones(nu), 1.0, 1.0
        self.EF, self.Ec, self.Esl, self.Ect = np.ones(L.n_fin), 1.0, 1.0, np.ones(nu)
        self.c = 1.0
        for _ in range(self.o.scaling):
            self._Dw_dev = self._dev(np.concatenate([self.Du, [self.Ds, self.Dt]]))
            self._launch(_SCALE)
            m = self._reduced(_SCALE, self._part[_SCALE][2]).cpu().numpy()
            Du, Ds, Dt, EF, Ec, Esl, Ect, c = self.Du, self.Ds, self.Dt, self.EF, self.Ec, self.Esl, self.Ect, self.c
            aF = np.abs(self.F) * EF[:, None] * Du[None, :]
            colP_u = c * np.max(np.abs(self.P_uu) * Du[:, None] * Du[None, :], axis=0)
            colA_u = np.maximum.reduce([aF.max(axis=0, initial=0.0), m[:nu], np.abs(self.ctrl) * Ect * Du])
            colA_s = max(abs(self.cvar_s) * Ec * Ds, m[nu], abs(self.sl) * Esl * Ds)
            colA_t = max(abs(self.cvar_t) * Ec * Dt, m[nu + 1])
            rowF = aF.max(axis=1, initial=0.0)
            rowc = max(abs(self.cvar_y) * Ec * m[nu + 2], abs(self.cvar_s) * Ec * Ds, abs(self.cvar_t) * Ec * Dt)
            rowsl = abs(self.sl) * Esl * Ds
            rowct = np.abs(self.ctrl) * Ect * Du
            self.Du = Du / np.sqrt(_guard(np.maximum(colP_u, colA_u)))
            self.Ds = Ds / float(np.sqrt(_guard(max(c * abs(self.P_ss) * Ds * Ds, colA_s))))
            self.Dt = Dt / float(np.sqrt(_guard(max(c * abs(self.P_tt) * Dt * Dt, colA_t))))
            self.EF = EF / np.sqrt(_guard(rowF))
            self.Ec = Ec / float(np.sqrt(_guard(rowc)))
            self.Esl = Esl / float(np.sqrt(_guard(rowsl)))
            self.Ect = Ect / np.sqrt(_guard(rowct))
            colP = c * np.concatenate([np.max(np.abs(self.P_uu) * self.Du[:, None] * self.Du[None, :], axis=0),
                                       [abs(self.P_ss) * self.Ds ** 2, abs(self.P_tt) * self.Dt ** 2]])
            qinf = c * np.max(np.abs(self.q_w * np.concatenate([self.Du, [self.Ds, self.Dt]])))
            g = 1.0 / max(colP.sum() / n, qinf, 1e-4)
            self.c = c * min(max(g, 1e-4), 1e4)
        self._Dw_dev = self._dev(np.concatenate([self.Du, [self.Ds, self.Dt]]))

    # -- factorisation ----------------------------------------------------------------------------
    def _global_rows(self):
        """Scaled bounds and rho of the sample-independent rows [final | cvar | slack | control]."""
        E = np.concatenate([self.EF, [self.Ec, self.Esl], self.Ect])
        ls, us = E * self.lo_g, E * self.hi_g
        ls[self.lo_g <= -_INF], us[self.hi_g >= _INF] = -_INF, _INF
        r = np.full(ls.shape, self.rho)
        r[np.abs(us - ls) < 1e-10] = 1e3 * self.rho
        r[(ls <= -_INF) & (us >= _INF)] = 1e-6
        return E, ls, us, r

    def _factor(self):
        L, o = self.L, self.o
        nu, nw, nf, nact = L.nu, L.nw, L.n_fin, L.nact
        nb = nact + 2
        self._launch(_GRAM)
        g = self._reduced(_GRAM, 0).cpu().numpy()
        idx = np.array(L.active + [nu, nu + 1])                  # pair index -> dense variable
        S = np.zeros((nw, nw))
        pr = np.array(L.pairs)
        a, bb = idx[pr[:, 0]], idx[pr[:, 1]]
        S[a, bb] += g[:L.npairs]
        off = a != bb
        S[bb[off], a[off]] += g[:L.npairs][off]
        h = np.zeros(nw); h[idx] = g[L.npairs:L.npairs + nb]
        eps_c = float(g[L.npairs + nb])
        ex = g[L.npairs + nb + 1:L.npairs + 2 * nb + 1]
        act = np.array(L.active)
        S[act, nu + 1] += ex[:nact]; S[nu + 1, act] += ex[:nact]
        S[nu, nu] += ex[nact]; S[nu + 1, nu + 1] += ex[nact + 1]
        E, ls, us, r = self._global_rows()
        rF, rc, rsl, rct = r[:nf], r[nf], r[nf + 1], r[nf + 2:]
        Fs = self.F * self.EF[:, None] * self.Du[None, :]
        ctS = self.Ect * self.ctrl * self.Du
        slS = self.Esl * self.sl * self.Ds
        vcw = np.array([self.Ec * self.cvar_s * self.Ds, self.Ec * self.cvar_t * self.Dt])
        S[:nu, :nu] += self.c * self.P_uu * self.Du[:, None] * self.Du[None, :] + Fs.T @ (rF[:, None] * Fs) + np.diag(rct * ctS ** 2)
        S[nu, nu] += self.c * self.P_ss * self.Ds ** 2 + rsl * slS ** 2
        S[nu + 1, nu + 1] += self.c * self.P_tt * self.Dt ** 2
        S += o.sigma * np.eye(nw)
        Sinv = np.linalg.inv(S)
        vfull = np.concatenate([np.zeros(nu), vcw])
        pw = Sinv @ (vfull - h)
        vp = float(vfull @ pw + eps_c - h @ pw)
        qw = self.c * self.q_w * np.concatenate([self.Du, [self.Ds, self.Dt]])
        self._h, self._vcw, self._Fs, self._ctS, self._slS, self._qw = h, vcw, Fs, ctS, slS, qw
        self._rg, self._lsg, self._usg, self._Eg = r, ls, us, E
        G, O = self.G, L.off
        put = lambda k, v: G[O[k]:O[k] + np.size(v)].copy_(self._dev(np.ravel(v)))
        put('Fs', Fs); put('ctS', ctS); put('slS', [slS]); put('vcw', vcw); put('rho', r); put('lo', ls); put('hi', us)
        put('Sinv', Sinv); put('h', h); put('pw', pw); put('scal', [eps_c, vp, o.sigma, o.alpha]); put('qw', qw)

    def reset(self, rho=None):
        """Forget the warm start (x, z, multipliers := 0) and optionally restart rho -- for a QP that has nothing in
        common with the previous one (the step from the relaxed first SCP iterations to the real constraints)."""
        for k in ("xy", "rloc", "zy", "ly", "zs", "ls"):
            self.st[k].zero_()
        O, L = self.L.off, self.L
        for k, n in (("z", L.ng), ("lam", L.ng), ("xw", L.nw), ("xt", L.nw + 1)):
            self.G[O[k]:O[k] + n].zero_()
        if rho is not None:
            self.rho = float(rho)
        self._stale = True

    # -- OSQP-style update ---------------------------------------------------------------------------
    def update(self, b=None):
        """New matrix values / bounds in the assembled buffers (same pattern, same scaling)."""
        if b is not None:
            self._b = b
        self._read_globals()
        self._factor()
        self._stale = True

    # -- solve --------------------------------------------------------------------------------------
    def _iterate(self, first):
        self._launch(_PASS, first)
        red = self._reduced(_PASS, 0)
        check(lib.saa_qp_dense_step(self.path.handle, self.G.data_ptr(), red.data_ptr(), int(first),
                                    self.path._stream()), self.path.handle)
        self.launches += 1

    def _residuals(self):
        L, o = self.L, self.o
        nu, nf, nw = L.nu, L.n_fin, L.nw
        O = L.off
        xin = torch.cat([self.G[O['xw']:O['xw'] + nw], self.G[O['lam'] + nf:O['lam'] + nf + 1]])
        self._launch(_CHECK, xin)
        r = self._reduced(_CHECK, 4).cpu().numpy()
        Gh = self.G.cpu().numpy()
        xw = Gh[O['xw']:O['xw'] + nw]
        zg, lg = Gh[O['z']:O['z'] + L.ng], Gh[O['lam']:O['lam'] + L.ng]
        u, s, t = xw[:nu], xw[nu], xw[nu + 1]
        # cvar row's product needs sum e_i y_i: the dense step's input 'cv' is for x~; recompute for x
        e_dot_y = self._cvar_dot()
        Ax_g = np.concatenate([self._Fs @ u, [self._vcw[0] * s + self._vcw[1] * t + e_dot_y], [self._slS * s], self._ctS * u])
        E = self._Eg
        rp = max(r[0], np.max(np.abs((Ax_g - zg) / E)))
        nAx = max(r[1], np.max(np.abs(Ax_g / E)))
        nz = max(r[2], np.max(np.abs(zg / E)))
        wF, wc, wsl, wct = lg[:nf], lg[nf], lg[nf + 1], lg[nf + 2:]
        gw = r[4:4 + nw].copy()
        gw[:nu] += self._Fs.T @ wF + self._ctS * wct
        gw[nu] += self._slS * wsl + wc * self._vcw[0]
        gw[nu + 1] += wc * self._vcw[1]
        Dw = np.concatenate([self.Du, [self.Ds, self.Dt]])
        Pw = self.c * np.concatenate([(self.P_uu * self.Du[:, None] * self.Du[None, :]) @ u,
                                      [self.P_ss * self.Ds ** 2 * s, self.P_tt * self.Dt ** 2 * t]])
        rd = max(np.max(np.abs((Pw + self._qw + gw) / Dw)), r[3]) / self.c
        ep = o.eps_abs + o.eps_rel * max(nAx, nz)
        ed = o.eps_abs + o.eps_rel * max(np.max(np.abs(Pw / Dw)), np.max(np.abs(gw / Dw)), r[3],
                                         np.max(np.abs(self._qw / Dw))) / self.c
        return rp, rd, ep, ed

    def _cvar_dot(self):
        """sum_i e_i y_i over all samples (CVaR row at the relaxed iterate x)."""
        v = (self.st['Dy'] * self.st['xy']).sum().reshape(1) * (self.Ec * self.cvar_y)
        if self._sharded:
            import torch.distributed as dist
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.group)
        return float(v.item())

    def solve(self):
        o = self.o
        t0 = time.perf_counter()
        status, it = 'maximum iterations reached', 0
        while it < o.max_iter:
            if self._stale:
                self._iterate(True)
                self._stale = False
            n = min(o.check_interval, o.max_iter - it)
            for _ in range(n):
                self._iterate(False)
            it += n
            rp, rd, ep, ed = self._residuals()
            if o.verbose:
                print(f"[device_qp] it {it} rp {rp:.3e}/{ep:.3e} rd {rd:.3e}/{ed:.3e} rho {self.rho:.3e}")
            if rp <= ep and rd <= ed:
                status = 'solved'
                break
            if o.adaptive_rho_interval and it % o.adaptive_rho_interval == 0:
                num, den = rp / max(ep, 1e-30), rd / max(ed, 1e-30)
                new_rho = float(np.clip(self.rho * np.sqrt(num / max(den, 1e-30)), 1e-6, 1e6))
                if new_rho > o.adaptive_rho_tolerance * self.rho or new_rho < self.rho / o.adaptive_rho_tolerance:
                    self.rho = new_rho
                    self._factor()
                    self._stale = True
                    self.refactorizations = getattr(self, 'refactorizations', 0) + 1
        L = self.L
        O = L.off
        xw = self.G[O['xw']:O['xw'] + L.nw].cpu().numpy()
        y_dev = self.st['Dy'] * self.st['xy']
        res = SimpleNamespace(u=self.Du * xw[:L.nu], slack=self.Ds * xw[L.nu], t=self.Dt * xw[L.nu + 1], y_dev=y_dev,
                              info=SimpleNamespace(status=status, iter=it, run_time=time.perf_counter() - t0))
        if not self._sharded:
            res.x = np.concatenate([res.u, y_dev.cpu().numpy(), [res.slack, res.t]])
        return res
