"""Sample sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

The reference is single-process; its only cross-sample operations are the sample
means of the final-state rows (``jnp.mean(..., axis=0)``, drone/drone_risk.py:294-296,
car/driving.py:311-313).  Everything else is independent per sample, so rank ``r``
owns the contiguous sample block ``shard_range(M, W, r)`` and with it a contiguous
sub-run of every sample-carrying CSC column.

Per SCP iteration:
  1. ``us`` (<= 60 doubles) is broadcast from rank 0;
  2. every rank runs ``saa_linearize_assemble`` on its block;
  3. ``all_reduce(sum)`` of the mean-row partial sums (123 / 120 doubles, NCCL);
  4. delivery of the row blocks, one of
     * ``'sharded'``  – blocks stay in their owners' HBM (compact matrices),
     * ``'peer'``     – fused gather: every rank's kernel stores straight into rank
       0's global value arrays through peer-mapped (CUDA IPC / NVLink) pointers,
     * ``'factored'`` – like ``'peer'`` but what crosses NVLink is the factored record
       (sensitivities + trajectory, 3.4 KB instead of 9.1 KB per sample); rank 0 expands it
       into the CSC entries with the same instructions a single GPU would use (drone only),
     * ``'nccl'``     – ``gather`` of the compact blocks to rank 0 followed by
       ``saa_merge_shard`` (device-to-device run copies).

``ShardedTailAssembler`` delivers the tail-reduced subproblem (``tail.py``) instead: every rank
keeps the ``K_r ~ (1 + margin) alpha M_r`` samples of ITS shard with the largest constraint
values (a stratified selection: the shards are i.i.d., so the global alpha-tail is inside the union
up to binomial fluctuations far below the margin; the a-posteriori check is the max over ranks of
``left_out_margin``) and stores their rows straight into rank 0's K-sample matrix over NVLink.
"""
import numpy as np
import torch
import torch.distributed as dist

from ._lib import lib, check


def shard_range(M, world, rank):
    """Balanced contiguous blocks: -> (first sample, sample count) of ``rank``."""
    M, world, rank = int(M), int(world), int(rank)
    if not (0 <= rank < world) or M < world:
        raise ValueError("need 0 <= rank < world <= M")
    base, rem = divmod(M, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def broadcast_controls(us_mat, src=0, group=None, device=None):
    """Rank ``src``'s iterate to every rank (host array in, host array out)."""
    t = torch.as_tensor(np.ascontiguousarray(us_mat, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy()


class PeerMeans:
    """All-reduce of the sample-mean sums fused with their finalisation, over NVLink peer memory.

    ``ncclAllReduce(123 f64)`` + ``saa_finalize_means`` cost ~75 us per step at N = 8 (collective
    latency + a second launch, serialised behind the 1.9 ms assemble kernel whose last block
    produces the sums).  ``saa_peer_allreduce_finalize`` is ONE launch: every rank stores its sums
    into every rank's inbox with peer stores, raises a flag, waits for the others' flags and adds the
    contributions in rank order -- a few microseconds, deterministic, bitwise identical on all ranks.
    (Running the means on a side stream instead does not work: the assemble kernel is a persistent
    grid that owns every register file, measured +0.16 ms.)"""

    def __init__(self, path, group=None):
        import ctypes as C
        self.path, self.group = path, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = int(lib.saa_peer_inbox_bytes())
        dev = path.device.index
        ptr, hd = C.c_void_p(), C.create_string_buffer(64)
        check(lib.saa_shared_alloc(dev, nbytes, C.byref(ptr), hd))
        self._own = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, hd.raw, group=group)
        self._ptrs = (C.c_void_p * self.world)()
        self._opened = []
        for q, raw in enumerate(handles):
            if q == self.rank:
                self._ptrs[q] = self._own
            else:
                p = C.c_void_p()
                check(lib.saa_shared_open(dev, raw, C.byref(p)))
                self._ptrs[q] = p.value
                self._opened.append(p.value)
        self.epoch = 0
        dist.barrier(group=group)

    def finalize(self, b, scp_iter=2):
        """After ``path.assemble(..., finalize=False)`` on the same stream, on EVERY rank."""
        p = self.path
        self.epoch += 1
        ptr = lambda t: t.data_ptr() if torch.is_tensor(t) else int(t)
        check(lib.saa_peer_allreduce_finalize(p.handle, p.mean_sums.data_ptr(), int(scp_iter), ptr(b['Ax']),
                                              ptr(b['l']), ptr(b['u']), self.rank, self.world, self._ptrs,
                                              self.epoch, p._stream()), p.handle)

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        dev = self.path.device.index
        for p in self._opened:
            lib.saa_shared_close(dev, p)
        self._opened = []
        dist.barrier(group=self.group)
        if self._own:
            lib.saa_shared_free(dev, self._own)
            self._own = None


def all_reduce_sums(t, group=None):
    """Sum of the per-rank partial sums (the ``mean`` is taken by ``saa_finalize_means``)."""
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class _DeviceArray:
    """Wraps a raw device pointer for ``torch.as_tensor`` (``__cuda_array_interface__``)."""

    def __init__(self, ptr, n, np_dtype):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr=np.dtype(np_dtype).str,
                                             data=(int(ptr), False), version=3)


class SharedBuffers:
    """Rank ``src`` owns cudaMalloc'ed (Ax, l, u); every other rank maps them over NVLink
    (CUDA IPC, opened with the mapping rank's GPU current)."""

    def __init__(self, sizes, np_dtype, device, src=0, group=None):
        import ctypes as C
        self.device, self.src = device, src
        self.rank = dist.get_rank(group)
        self.owner = self.rank == src
        esz = np.dtype(np_dtype).itemsize
        self.ptrs, handles = [], []
        if self.owner:
            for n in sizes:
                ptr, hd = C.c_void_p(), C.create_string_buffer(64)
                check(lib.saa_shared_alloc(device.index, int(n) * esz, C.byref(ptr), hd))
                self.ptrs.append(ptr.value)
                handles.append(hd.raw)
        payload = [handles]
        dist.broadcast_object_list(payload, src=src, group=group)
        if not self.owner:
            for hd in payload[0]:
                ptr = C.c_void_p()
                check(lib.saa_shared_open(device.index, hd, C.byref(ptr)))
                self.ptrs.append(ptr.value)
        self.tensors = None
        if self.owner:
            self.tensors = [torch.as_tensor(_DeviceArray(p, n, np_dtype), device=device)
                            for p, n in zip(self.ptrs, sizes)]

    def close(self):
        for p in self.ptrs:
            (lib.saa_shared_free if self.owner else lib.saa_shared_close)(self.device.index, p)
        self.ptrs = []


class ShardedAssembler:
    """Drives one ``DevicePath`` per rank.  ``path`` must have been created with
    ``M_local, M_global, sample_offset = shard_range(...)``-consistent arguments."""

    def __init__(self, path, mode='sharded', group=None):
        if mode not in ('sharded', 'peer', 'factored', 'nccl'):
            raise ValueError("mode must be 'sharded', 'peer', 'factored' or 'nccl'")
        self.path, self.mode, self.group = path, mode, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.out = None
        if mode in ('sharded', 'nccl'):
            path.set_output_geometry(path.M_local, 0)
        else:
            path.set_output_geometry(path.M_global, path.sample_offset)
        if mode in ('peer', 'factored'):
            self._setup_peer()
        if mode == 'nccl':
            self._setup_nccl()

    # -- fused gather: rank 0 owns the global arrays, everybody maps them -------------
    def _setup_peer(self):
        p = self.path
        n_rows, _, nnz = p.pattern_sizes(False)
        npdt = np.float64 if p.bits == 64 else np.float32
        sizes = [nnz, n_rows, n_rows]
        if self.mode == 'factored':
            sizes += list(p.factored_sizes())
            counts = [None] * self.world
            dist.all_gather_object(counts, p.M_local, group=self.group)
            self.counts = counts
        self.shared = SharedBuffers(sizes, npdt, p.device, 0, self.group)
        # rank 0: tensors over its own allocations; other ranks: raw peer pointers into rank 0's HBM
        bufs = self.shared.tensors if self.rank == 0 else self.shared.ptrs
        self.out = dict(Ax=bufs[0], l=bufs[1], u=bufs[2], const_state=None)
        if self.mode == 'factored':
            self.fsp, self.fp = bufs[3], bufs[4]

    def _setup_nccl(self):
        p = self.path
        counts = [torch.zeros(1, dtype=torch.int64, device=p.device) for _ in range(self.world)]
        dist.all_gather(counts, torch.tensor([p.M_local], dtype=torch.int64, device=p.device),
                        group=self.group)
        self.counts = [int(c.item()) for c in counts]
        if len(set(self.counts)) != 1:
            raise ValueError("mode='nccl' gathers equal shards: M must be divisible by the world size")
        if self.rank == 0:
            from .device_path import DevicePath
            # a second handle on rank 0 describing the GLOBAL matrix: pattern, constant
            # entries, merge offsets, mean rows.  It needs the problem constants
            # (bind_global_params) but no samples.
            self.global_path = DevicePath(p.problem, p.method, p.S, p.alpha, p.M_global,
                                          M_global=p.M_global, sample_offset=0, variant=p.variant,
                                          precision='fp64' if p.bits == 64 else 'fp32',
                                          device=p.device.index)

    def bind_global_params(self, setter):
        """``setter(path)`` sets the problem constants (``set_params_drone`` / ``_car``) on rank
        0's global-geometry handle; call once after construction in ``mode='nccl'``."""
        if self.mode == 'nccl' and self.rank == 0:
            setter(self.global_path)

    # -- one SCP iteration ---------------------------------------------------------------
    def step(self, us_mat, scp_iter):
        p = self.path
        if p._uses_relaxed_pattern(scp_iter):
            # car, scp_iter 0 (car/driving.py:411-415): every sample row is multiplied by 0, so there
            # are no row blocks to shard or gather -- the O(1)-sized problem is assembled by one rank.
            # The buffers and merge offsets of this class are those of the normal pattern.
            raise ValueError("car scp_iter 0 has no sample rows to shard: assemble it on one rank "
                             "(Model.get_constraints_coeffs / DevicePath.assemble)")
        us = broadcast_controls(us_mat, 0, self.group, device=p.device if dist.get_backend(self.group) == 'nccl' else None)
        if self.mode == 'factored' and self.rank != 0:
            # constants of this rank's sample slice (remote, once per relaxation state), then the
            # factored record of the block instead of its CSC entries
            p.write_constants(self.out, scp_iter, write_shared=False)
            p.linearize_factored(us, scp_iter, self.fsp, self.fp, self.out['u'])
            b = self.out
        else:
            remote = self.mode in ('peer', 'factored')
            b = p.assemble(us, scp_iter, finalize=False, write_shared=(self.rank == 0 or not remote),
                           out=self.out)
        all_reduce_sums(p.mean_sums, self.group)
        if self.mode == 'factored':
            torch.cuda.current_stream(p.device).synchronize()
            dist.barrier(group=self.group)            # every record has landed in rank 0's HBM
            if self.rank == 0:
                p.expand_factored(scp_iter, self.fsp, self.fp, self.counts[0], p.M_global - self.counts[0],
                                  b['Ax'])
                p.finalize_means(b, scp_iter)
            return b if self.rank == 0 else None
        if self.mode == 'peer':
            # remote stores must have landed before rank 0 consumes the arrays
            torch.cuda.current_stream(p.device).synchronize()
            dist.barrier(group=self.group)
            if self.rank == 0:
                p.finalize_means(b, scp_iter)
            return b if self.rank == 0 else None
        p.finalize_means(b, scp_iter)
        if self.mode == 'sharded':
            return b
        return self._gather_nccl(b, scp_iter)

    def _gather_nccl(self, b, scp_iter):
        p = self.path
        n_rows, _, nnz = p.pattern_sizes(False)
        n_var = self._n_var(p)
        sendA, sendu = b['Ax'][:n_var], b['u']
        if self.rank == 0:
            recvA = [torch.empty_like(sendA) for _ in range(self.world)]
            recvu = [torch.empty_like(sendu) for _ in range(self.world)]
        else:
            recvA = recvu = None
        dist.gather(sendA, recvA, dst=0, group=self.group)
        dist.gather(sendu, recvu, dst=0, group=self.group)
        if self.rank != 0:
            return None
        g = self.global_path
        gb = g.buffers(False)
        relaxed = scp_iter < g.relax_threshold()
        if gb.get('const_state') != relaxed:
            # rank 0 writes every constant entry of the global matrix (once per relaxation state)
            check(lib.saa_write_constants(g.handle, int(scp_iter), 1, gb['Ax'].data_ptr(),
                                          gb['l'].data_ptr(), gb['u'].data_ptr(), g._stream()), g.handle)
            gb['const_state'] = relaxed
        first = 0
        for r in range(self.world):
            check(lib.saa_merge_shard(g.handle, recvA[r].data_ptr(), recvu[r].data_ptr(), self.counts[r],
                                      first, gb['Ax'].data_ptr(), gb['u'].data_ptr(), g._stream()), g.handle)
            first += self.counts[r]
        g.mean_sums.copy_(p.mean_sums)
        g.finalize_means(gb, scp_iter)
        return gb

    @staticmethod
    def _n_var(p):
        """Length of the leading u-column block of A.data (everything before the y columns):
        per sample 1140 (drone) / 380 (car) Jacobian entries, plus the final-row and
        control-row entries (117 + 60 / 116 + 40)."""
        from . import _lib
        if p.problem == _lib.SAA_DRONE:
            return 1140 * p.M_out + 177
        return 380 * p.M_out + 156


def tie_quotas(gt_eq, K_total):
    """Per-rank (count, ties taken) of an exact global top-``K_total`` selection, from every rank's
    ``(#{Z > threshold}, #{Z == threshold})``: all values above the threshold are taken, the remaining places go to
    the ties in rank order (= global index order, the tie rule of the single-GPU selection)."""
    gt_eq = np.asarray(gt_eq, dtype=np.int64).reshape(-1, 2)
    rem = int(K_total) - int(gt_eq[:, 0].sum())
    if rem < 0 or rem > int(gt_eq[:, 1].sum()):
        raise ValueError("inconsistent threshold counts for the requested K")
    out = []
    for gt, eq in gt_eq:
        take = int(min(eq, rem))
        rem -= take
        out.append((int(gt) + take, take))
    return out


def exact_global_select(path, Z, K_total, idx_out, group=None):
    """Exact global top-``K_total`` of the sharded values ``Z`` (this rank's ``path.M_local`` of them):
    the library's radix select with its histogram all-reduced between the passes.  Writes this rank's
    selected local indices (ascending) to ``idx_out`` and returns the list of per-rank counts (ties at the
    threshold go to the lower rank = lower global index, as in the single-GPU selection)."""
    h = path.handle
    st = path._stream()
    hist = torch.zeros(256, dtype=torch.int32, device=path.device)
    check(lib.saa_select_begin(h, int(K_total), st), h)
    for p in range(int(lib.saa_select_passes(h))):
        check(lib.saa_select_pass_hist(h, Z.data_ptr(), p, hist.data_ptr(), st), h)
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        check(lib.saa_select_pass_pick(h, hist.data_ptr(), p, st), h)
    cnt = torch.zeros(2, dtype=torch.int64, device=path.device)
    check(lib.saa_select_counts(h, Z.data_ptr(), cnt.data_ptr(), st), h)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    allc = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt, group=group)
    allc = torch.stack(allc).cpu().numpy()                  # the one host synchronisation of the selection
    quotas = tie_quotas(allc, K_total)
    check(lib.saa_select_finish(h, Z.data_ptr(), quotas[rank][1], idx_out.data_ptr(), st), h)
    return [c for c, _ in quotas]


class ShardedTailAssembler:
    """Tail-reduced subproblem of a sharded sample set, assembled on rank 0 (peer stores).

    ``path``: this rank's ``DevicePath`` over its shard (``M_local`` of ``M_global`` samples, params
    and samples set).  Rank r contributes its ``K_r`` worst samples as samples
    ``offset_r .. offset_r + K_r`` of a matrix for ``K_total = sum K_r`` samples; the CVaR row keeps
    ``M_global alpha t`` and the expectation rows the mean over all ``M_global`` samples."""

    def __init__(self, path, margin=0.25, K_local=None, mode='peer', group=None, exact=False):
        """``mode='peer'``: the ranks' kernels store the CSC entries of their selected samples into
        rank 0's arrays; ``mode='factored'`` (drone): they store the factored record (3.4 instead of
        9.1 KB per sample cross NVLink) and rank 0 expands it, bitwise identically.
        ``exact=True``: the K_total samples with the largest Z_i of the WHOLE set (``exact_global_select``:
        all-reduced radix histograms) instead of each rank's own top K_r; the number a rank contributes
        then changes from iteration to iteration (``self.counts``), the matrix on rank 0 is the
        single-GPU tail-reduced matrix of all samples, sample for sample."""
        from .tail import TailSubproblem
        from . import _lib
        if mode not in ('peer', 'factored'):
            raise ValueError("mode must be 'peer' or 'factored'")
        if mode == 'factored' and path.problem != _lib.SAA_DRONE:
            raise ValueError("the factored record exists for the drone only")
        self.path, self.group, self.mode = path, group, mode
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        K = int(min(path.M_local, max(1, int(np.ceil((1.0 + margin) * path.alpha * path.M_local))))
                if K_local is None else K_local)
        counts = [None] * self.world
        dist.all_gather_object(counts, K, group=group)
        self.counts = [int(c) for c in counts]
        self.K_total, self.offset = sum(self.counts), sum(self.counts[:self.rank])
        self.exact = bool(exact)
        if self.exact:
            # any rank may hold the whole tail: capacity = min(its shard, K_total)
            K = min(path.M_local, self.K_total)
        self.tail = TailSubproblem(path, K=K, out_geometry=(self.K_total, self.offset if not exact else 0))
        sub = self.tail.sub
        n_rows, _, nnz = sub.pattern_sizes(False)
        npdt = np.float64 if sub.bits == 64 else np.float32
        sizes = [nnz, n_rows, n_rows] + (list(sub.factored_sizes()) if mode == 'factored' else [])
        self.shared = SharedBuffers(sizes, npdt, sub.device, 0, group)
        bufs = self.shared.tensors if self.rank == 0 else self.shared.ptrs
        self.out = dict(Ax=bufs[0], l=bufs[1], u=bufs[2], const_state=None)
        if mode == 'factored':
            self.fsp, self.fp = bufs[3], bufs[4]
        kmax = self.K_total if self.exact else max(self.counts)  # all_gather wants equal sizes: pad
        self._idx_send = torch.zeros(kmax, dtype=torch.int64, device=sub.device)
        self._idx_all = [torch.empty(kmax, dtype=torch.int64, device=sub.device) for _ in self.counts]

    def pattern(self):
        """(n_rows, n_cols, indptr, indices) of the K_total-sample matrix (rank 0's result)."""
        return self.tail.sub.pattern(False)

    def step(self, us_mat, scp_iter):
        """One SCP iteration.  Rank 0 gets (dict(Ax, l, u) of device tensors, idx: the selected
        samples' GLOBAL indices in the order of the matrix's sample blocks); other ranks (None, None)."""
        p, t = self.path, self.tail
        if scp_iter < p.relax_threshold() and p._uses_relaxed_pattern(scp_iter):
            raise ValueError("the sharded tail gather covers scp_iter >= 1 for the car")
        nccl = dist.get_backend(self.group) == 'nccl'
        us = broadcast_controls(us_mat, 0, self.group, device=p.device if nccl else None)
        if self.exact:
            return self._step_exact(us, scp_iter)
        if self.mode == 'factored' and self.rank != 0:
            us = t.select(us)
            t.sub.write_constants(self.out, scp_iter, write_shared=False)
            t.sub.linearize_factored(us, scp_iter, self.fsp, self.fp, self.out['u'])
            b = self.out
        else:
            b = t.assemble(us, scp_iter, out=self.out, write_shared=(self.rank == 0), finalize=False)
        all_reduce_sums(p.mean_sums, self.group)                 # sums over ALL samples of all ranks
        self._idx_send[:t.K] = t.idx + p.sample_offset
        dist.all_gather(self._idx_all, self._idx_send, group=self.group)
        torch.cuda.current_stream(p.device).synchronize()
        dist.barrier(group=self.group)                           # remote rows have landed in rank 0's HBM
        if self.rank != 0:
            return None, None
        if self.mode == 'factored' and self.K_total > self.counts[0]:
            t.sub.expand_factored(scp_iter, self.fsp, self.fp, self.counts[0], self.K_total - self.counts[0],
                                  b['Ax'])
        t.finalize_means(b, scp_iter)
        return b, torch.cat([v[:c] for v, c in zip(self._idx_all, self.counts)])

    def _step_exact(self, us, scp_iter):
        p, t = self.path, self.tail
        f, s = t.full, t.sub
        st = f._stream()
        usc = f._us(us)
        check(lib.saa_linearize_means(f.handle, usc.ctypes.data, t.Z.data_ptr(), f.mean_sums.data_ptr(), st), f.handle)
        self.counts = exact_global_select(f, t.Z, self.K_total, t.idx, self.group)
        K_r, off = self.counts[self.rank], sum(self.counts[:self.rank])
        t.K = K_r
        s.set_active(K_r, self.K_total, off)
        check(lib.saa_gather_samples(s.handle, f.handle, t.idx.data_ptr(), st), s.handle)
        # the slice [off, off + K_r) moves from iteration to iteration: its constants are rewritten each time
        # (the slices of all ranks tile the K_total samples, so every constant entry is covered)
        self.out['const_state'] = None
        if self.mode == 'factored' and self.rank != 0:
            s.write_constants(self.out, scp_iter, write_shared=False)
            s.linearize_factored(usc, scp_iter, self.fsp, self.fp, self.out['u'])
        else:
            s.assemble(usc, scp_iter, out=self.out, write_shared=(self.rank == 0), finalize=False)
        all_reduce_sums(p.mean_sums, self.group)
        self._idx_send[:K_r] = t.idx[:K_r] + p.sample_offset
        dist.all_gather(self._idx_all, self._idx_send, group=self.group)
        torch.cuda.current_stream(p.device).synchronize()
        dist.barrier(group=self.group)
        if self.rank != 0:
            return None, None
        b = self.out
        if self.mode == 'factored' and self.K_total > self.counts[0]:
            s.expand_factored(scp_iter, self.fsp, self.fp, self.counts[0], self.K_total - self.counts[0], b['Ax'])
        t.finalize_means(b, scp_iter)
        return b, torch.cat([v[:c] for v, c in zip(self._idx_all, self.counts)])

    def left_out_margin(self, t_risk):
        """max over all ranks of ``TailSubproblem.left_out_margin``."""
        v = torch.tensor([self.tail.left_out_margin(t_risk)], dtype=torch.float64, device=self.path.device)
        dist.all_reduce(v, op=dist.ReduceOp.MAX, group=self.group)
        return float(v.item())

    def close(self):
        self.shared.close()


# --------------------------------------------------------------------------------------------
# In-run parity self-check of the multi-GPU delivery modes (bench.py, N > 1)
# --------------------------------------------------------------------------------------------
def _rel_err(a, b):
    """max |a - b| / max(|b|, 1e-300) over finite entries; inf if the non-finite patterns differ."""
    a, b = a.double(), b.double()
    fin = torch.isfinite(b)
    if not torch.equal(fin, torch.isfinite(a)) or not torch.equal(a[~fin], b[~fin]):
        return float('inf')
    if not bool(fin.any()):
        return 0.0
    return float(((a[fin] - b[fin]).abs() / b[fin].abs().clamp_min(1e-300)).max().item())


def _sample_runs(Ax, indptr, n_fin_rows, M_mat, sample_pos, nu, indices):
    """u-column entries of the samples at positions ``sample_pos`` of a matrix with ``M_mat`` samples,
    concatenated column by column (device tensor)."""
    out = []
    for c in range(nu):
        lo, hi = int(indptr[c]), int(indptr[c + 1])
        nf = int((indices[lo:min(hi, lo + 4)] < n_fin_rows).sum())
        L = (hi - lo - nf - 1) // M_mat
        if L == 0:
            continue
        base = lo + nf
        idx = (base + sample_pos[:, None] * L + torch.arange(L, device=Ax.device)[None, :]).reshape(-1)
        out.append(Ax[idx])
    return torch.cat(out)


def parity_self_check(make_path, set_params, M, us, scp_iter=2, group=None, modes=None):
    """Every delivery mode on ``world`` ranks against ONE GPU assembling all ``M`` samples.

    ``make_path(first, count, M_global) -> DevicePath`` with params and samples set (every rank must be
    able to build any shard, i.e. the samples are generated identically on all ranks).
    -> {mode: max relative error} (identical on every rank).  GPUTEST runs on one GPU; this makes the
    multi-GPU gathers verifiable in the driver's own scaling run."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, cnt = shard_range(M, world, rank)
    single = make_path(0, M, M)
    dev = single.device
    ref = {k: v.clone() for k, v in single.assemble(us, scp_iter).items() if torch.is_tensor(v)}
    n_rows_g, _, indptr_g, indices_g = single.pattern(False)
    nu = single.S * (3 if single.problem == 0 else 2)
    n_fin = 6 if single.problem == 0 else 4
    R = single.S * (3 if single.problem == 0 else 1)
    row_s0_g = n_fin + 1 + M
    res = {}
    modes = modes or (('sharded', 'sharded_peer_means', 'peer', 'nccl') + (('factored',) if single.problem == 0 else ())
                      + ('tail', 'tail_exact'))
    for mode in modes:
        err = 0.0
        if mode == 'nccl' and M % world:
            continue
        if mode in ('tail', 'tail_exact'):
            asm = ShardedTailAssembler(make_path(first, cnt, M), margin=0.5,
                                       mode='factored' if single.problem == 0 else 'peer', group=group,
                                       exact=(mode == 'tail_exact'))
            for _ in range(2 if asm.exact else 1):       # twice: the active counts are set again
                b, idx = asm.step(us, scp_iter)
            if rank == 0 and asm.exact:
                # the single-GPU tail-reduced matrix of ALL samples, entry for entry
                from .tail import TailSubproblem
                ts = TailSubproblem(single, K=asm.K_total)
                bs = ts.assemble(us, scp_iter)
                torch.cuda.synchronize()
                if not torch.equal(ts.idx, idx):
                    err = float('inf')
                for k in ('Ax', 'l', 'u'):
                    err = max(err, _rel_err(b[k], bs[k]))
                del ts
            if rank == 0:
                torch.cuda.synchronize()
                n_rows_t, _, indptr_t, indices_t = asm.pattern()
                K = idx.numel()
                got = _sample_runs(b['Ax'], indptr_t, n_fin, K, torch.arange(K, device=dev), nu, indices_t)
                want = _sample_runs(ref['Ax'], indptr_g, n_fin, M, idx.to(dev), nu, indices_g)
                err = max(err, _rel_err(got, want))
                rows_t = (n_fin + 1 + K + torch.arange(K * R, device=dev))
                rows_g = (row_s0_g + (idx.to(dev)[:, None] * R + torch.arange(R, device=dev)[None, :]).reshape(-1))
                err = max(err, _rel_err(b['u'][rows_t], ref['u'][rows_g]))
                err = max(err, _rel_err(b['u'][:n_fin], ref['u'][:n_fin]), _rel_err(b['l'][:n_fin], ref['l'][:n_fin]))
            asm.close()
        elif mode == 'sharded_peer_means':
            # row blocks stay sharded; the expectation rows come from the fused peer-memory all-reduce
            path = make_path(first, cnt, M)
            path.set_output_geometry(cnt, 0)
            pm = PeerMeans(path, group=group)
            for _ in range(3):                           # several epochs: flags / double buffer
                b = path.assemble(us, scp_iter, finalize=False)
                pm.finalize(b, scp_iter)
            torch.cuda.synchronize()
            err = max(err, _rel_err(b['u'][:n_fin], ref['u'][:n_fin]), _rel_err(b['l'][:n_fin], ref['l'][:n_fin]))
            _, _, indptr_s, indices_s = path.pattern(False)
            fin_s = torch.as_tensor(np.flatnonzero(indices_s[:int(indptr_s[nu])] < n_fin), device=dev)
            fin_g = torch.as_tensor(np.flatnonzero(indices_g[:int(indptr_g[nu])] < n_fin), device=dev)
            err = max(err, _rel_err(b['Ax'][fin_s], ref['Ax'][fin_g]))
            # bitwise identical on every rank (same numbers added in the same order)
            mine = torch.cat([b['l'][:n_fin].double(), b['Ax'][fin_s].double()])
            lo, hi = mine.clone(), mine.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
            if not torch.equal(lo, hi):
                err = float('inf')
            pm.close()
            del path
        else:
            path = make_path(first, cnt, M)
            asm = ShardedAssembler(path, mode=mode, group=group)
            asm.bind_global_params(set_params)
            b = asm.step(us, scp_iter)
            torch.cuda.synchronize()
            dist.barrier(group=group)
            if mode == 'sharded':
                _, _, indptr_s, indices_s = path.pattern(False)
                got = _sample_runs(b['Ax'], indptr_s, n_fin, cnt, torch.arange(cnt, device=dev), nu, indices_s)
                want = _sample_runs(ref['Ax'], indptr_g, n_fin, M, first + torch.arange(cnt, device=dev), nu, indices_g)
                err = max(err, _rel_err(got, want))
                rs = n_fin + 1 + cnt
                err = max(err, _rel_err(b['u'][rs:rs + cnt * R], ref['u'][row_s0_g + first * R:row_s0_g + (first + cnt) * R]))
                err = max(err, _rel_err(b['u'][:n_fin], ref['u'][:n_fin]), _rel_err(b['l'][:n_fin], ref['l'][:n_fin]))
            elif rank == 0:
                for k in ('Ax', 'l', 'u'):
                    err = max(err, _rel_err(b[k], ref[k]))
            if mode in ('peer', 'factored'):
                asm.shared.close()
            del asm, path
        t = torch.tensor([err], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        res[mode] = float(t.item())
        torch.cuda.synchronize()
        dist.barrier(group=group)
    return res
