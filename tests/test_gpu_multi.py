"""2-rank NCCL runs of the sharded path (skipped on a single-GPU box): every delivery
mode must reproduce the single-GPU matrix."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, initfile, M, problem, out):
    import torch
    import torch.distributed as dist
    from riskaversetrajopt_b200 import _lib, dist as sd
    from riskaversetrajopt_b200.device_path import DevicePath
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{initfile}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    rs = np.random.RandomState(0)
    if problem == 'drone':
        from riskaversetrajopt_b200.drone import drone_params as dp
        from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
        np.random.seed(0)
        DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)
        us = rs.randn(20, 3)
        pid, its = _lib.SAA_DRONE, (0, 2, 3)

        def make(first, cnt, M_global):
            p = DevicePath(pid, 'saa', 20, 0.1, cnt, M_global=M_global, sample_offset=first, device=rank)
            p.set_params_drone(dp, dp.OSQP_TOL)
            p.set_samples_drone(masses[first:first + cnt], DWs[first:first + cnt], obs_Qs[first:first + cnt])
            return p
        setp = lambda p: p.set_params_drone(dp, dp.OSQP_TOL)
    else:
        from riskaversetrajopt_b200.car import driving_params as cp
        from riskaversetrajopt_b200.car.driving import sample_uncertain_parameters, BETA
        np.random.seed(0)
        s = sample_uncertain_parameters(M, 'saa')
        us = 0.01 + 0.3 * rs.randn(20, 2)
        pid, its = _lib.SAA_CAR, (1, 2)

        def make(first, cnt, M_global):
            p = DevicePath(pid, 'saa', 20, 0.05, cnt, M_global=M_global, sample_offset=first, device=rank)
            p.set_params_car(cp, BETA, cp.OSQP_TOL)
            p.set_samples_car(*(x[first:first + cnt] for x in s))
            return p
        setp = lambda p: p.set_params_car(cp, BETA, cp.OSQP_TOL)

    res = {}
    first, cnt = sd.shard_range(M, world, rank)
    for mode in ('sharded', 'peer', 'nccl') + (('factored',) if problem == 'drone' else ()):
        if mode == 'nccl' and M % world:
            continue            # the NCCL gather needs equal shards (ValueError otherwise)
        path = make(first, cnt, M)
        asm = sd.ShardedAssembler(path, mode=mode)
        asm.bind_global_params(setp)
        for it in its:
            b = asm.step(us if rank == 0 else np.zeros_like(us), it)
            torch.cuda.synchronize()
            dist.barrier()
            if b is not None:
                res[(mode, it, rank)] = tuple(b[k].cpu().numpy() for k in ('Ax', 'l', 'u'))
        del asm, path
    if rank == 0:
        single = make(0, M, M)
        for it in its:
            b = single.assemble(us, it)
            res[('single', it, 0)] = tuple(b[k].cpu().numpy() for k in ('Ax', 'l', 'u'))
        if problem == 'drone':
            res['pattern'] = single.pattern()[2]
    out[rank] = res
    dist.destroy_process_group()


@pytest.mark.parametrize("problem,M", [("drone", 1000), ("car", 1000), ("drone", 37)])
def test_two_ranks_reproduce_single_gpu(problem, M):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    mgr = mp.Manager()
    out = mgr.dict()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, os.path.join(d, "init"), M, problem, out), nprocs=2, join=True)
    r0, r1 = out[0], out[1]
    its = (0, 2, 3) if problem == 'drone' else (1, 2)
    per, fixed = (1140, 177) if problem == 'drone' else (380, 156)
    R = 60 if problem == 'drone' else 20
    nfin = 6 if problem == 'drone' else 4
    for it in its:
        Ax, l, u = r0[('single', it, 0)]
        for mode in ('peer', 'nccl', 'factored'):
            if (mode == 'nccl' and M % 2) or (mode == 'factored' and problem != 'drone'):
                continue
            gAx, gl, gu = r0[(mode, it, 0)]
            if mode == 'factored':
                # same instructions form the products on rank 0: bitwise equal to the single-GPU run
                body_cols = slice(0, per * M + fixed)
                diff = np.flatnonzero(gAx[body_cols] != Ax[body_cols])
                # only the <= 117 mean-row entries may differ (summation order across ranks)
                assert diff.size <= 117, (mode, it, diff.size)
            # u-column block, bounds and the mean rows: same kernel, same samples -> bitwise, except
            # the mean rows (different summation order across ranks)
            assert np.allclose(gAx, Ax, rtol=1e-12, atol=1e-14), (mode, it)
            body = slice(nfin, None)
            assert np.array_equal(gu[body], u[body]) and np.array_equal(gl[body], l[body])
            assert np.allclose(gu[:nfin], u[:nfin], rtol=1e-12, atol=1e-14)
        # sharded: each rank's compact block equals the corresponding rows of the global problem
        M0 = (M + 1) // 2
        for rank, res, first, cnt in ((0, r0, 0, M0), (1, r1, M0, M - M0)):
            sAx, sl, su = res[('sharded', it, rank)]
            assert sAx.size == (per + (R + 2) + 1 + R) * cnt + fixed + 3
            row_s0_g, row_s0_s = nfin + 1 + M, nfin + 1 + cnt
            assert np.array_equal(su[row_s0_s:row_s0_s + cnt * R], u[row_s0_g + first * R:row_s0_g + (first + cnt) * R])
            assert np.allclose(su[:nfin], u[:nfin], rtol=1e-12, atol=1e-14)      # global means on every rank


def _tail_worker(rank, world, initfile, M, problem, mode, out, exact=False, skew=None):
    import torch
    import torch.distributed as dist
    from riskaversetrajopt_b200 import _lib, dist as sd
    from riskaversetrajopt_b200.device_path import DevicePath
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{initfile}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    rs = np.random.RandomState(0)
    if problem == 'drone':
        from riskaversetrajopt_b200.drone import drone_params as dp
        from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
        np.random.seed(0)
        DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)
        us, its, alpha = 0.1 * rs.randn(20, 3), (0, 2), 0.1
        if skew:
            # all the worst samples on one rank: the other contributes (next to) nothing to an exact tail
            p = DevicePath(_lib.SAA_DRONE, 'saa', 20, alpha, M, device=rank)
            p.set_params_drone(dp, dp.OSQP_TOL); p.set_samples_drone(masses, DWs, obs_Qs)
            Z = torch.empty(M, dtype=torch.float64, device=f'cuda:{rank}')
            p.assemble(us, 2, Z=Z)
            order = np.argsort(-Z.cpu().numpy(), kind='stable')
            order = order if skew == 'first' else order[::-1].copy()
            DWs, masses, obs_Qs = DWs[order], masses[order], obs_Qs[order]
            p.close()

        def make(first, cnt):
            p = DevicePath(_lib.SAA_DRONE, 'saa', 20, alpha, cnt, M_global=M, sample_offset=first, device=rank)
            p.set_params_drone(dp, dp.OSQP_TOL)
            p.set_samples_drone(masses[first:first + cnt], DWs[first:first + cnt], obs_Qs[first:first + cnt])
            return p
    else:
        from riskaversetrajopt_b200.car import driving_params as cp
        from riskaversetrajopt_b200.car.driving import sample_uncertain_parameters, BETA
        np.random.seed(0)
        s = sample_uncertain_parameters(M, 'saa')
        us, its, alpha = 0.01 + 0.3 * rs.randn(20, 2), (1, 2), 0.1

        def make(first, cnt):
            p = DevicePath(_lib.SAA_CAR, 'saa', 20, alpha, cnt, M_global=M, sample_offset=first, device=rank)
            p.set_params_car(cp, BETA, cp.OSQP_TOL)
            p.set_samples_car(*(x[first:first + cnt] for x in s))
            return p
    res = {}
    first, cnt = sd.shard_range(M, world, rank)
    asm = sd.ShardedTailAssembler(make(first, cnt), margin=0.5, mode=mode, exact=exact)
    for it in its:
        b, idx = asm.step(us if rank == 0 else np.zeros_like(us), it)
        Zloc = asm.tail.Z.cpu().numpy()
        sel = asm.tail.idx[:asm.tail.K].cpu().numpy()
        res[('local', it)] = (Zloc, sel, first)
        if rank == 0:
            res[('tail', it)] = tuple(b[k].cpu().numpy() for k in ('Ax', 'l', 'u')) + (idx.cpu().numpy(),)
    res['margin'] = asm.left_out_margin(1e9)
    if rank == 0:
        res['pattern'] = asm.pattern()
        single = make(0, M)
        for it in its:
            b = single.assemble(us, it)
            res[('single', it)] = tuple(b[k].cpu().numpy() for k in ('Ax', 'l', 'u'))
        res['single_pattern'] = single.pattern()
    out[rank] = res
    asm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("problem,M,mode,exact,skew", [
    ("drone", 1000, "peer", False, None), ("car", 600, "peer", False, None), ("drone", 77, "peer", False, None),
    ("drone", 1000, "factored", False, None), ("drone", 77, "factored", False, None),
    ("drone", 1000, "peer", True, None), ("car", 600, "peer", True, None), ("drone", 1000, "factored", True, None),
    ("drone", 400, "peer", True, "first"), ("drone", 400, "peer", True, "last"),
    ("drone", 400, "factored", True, "first"), ("drone", 400, "factored", True, "last")])
def test_two_ranks_tail_subproblem(problem, M, mode, exact, skew):
    """Tail selection on 2 ranks -- stratified (each rank's own top K_r) or exact (the global top K_total,
    whatever the split; ``skew`` puts all of the tail on one rank) -- rows stored into rank 0's K-sample
    matrix over NVLink: the result is the single-GPU full matrix restricted to the selected samples."""
    import scipy.sparse as sp
    import torch
    import torch.multiprocessing as mp
    from test_gpu_tail import _select_ref, _submatrix
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    mgr = mp.Manager()
    out = mgr.dict()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_tail_worker, args=(2, os.path.join(d, "init"), M, problem, mode, out, exact, skew), nprocs=2, join=True)
    r0, r1 = out[0], out[1]
    its = (0, 2) if problem == 'drone' else (1, 2)
    R, nu, nfin = (60, 60, 6) if problem == 'drone' else (20, 40, 4)
    n_rows, n_cols, indptr, indices = r0['pattern']
    fr, fc, findptr, findices = r0['single_pattern']
    assert r0['margin'] < 0 and r1['margin'] == r0['margin']
    for it in its:
        Ax, l, u, idx = r0[('tail', it)]
        # every rank kept the K_r largest of its own shard; rank 0 holds their global indices
        parts = []
        for r in (r0, r1):
            Zloc, sel, first = r[('local', it)]
            assert np.array_equal(sel, _select_ref(Zloc, len(sel)))
            parts.append(sel + first)
        assert np.array_equal(idx, np.concatenate(parts))
        if exact:
            Zall = np.concatenate([r0[('local', it)][0], r1[('local', it)][0]])
            assert np.array_equal(idx, _select_ref(Zall, len(idx)))       # the global top K_total
            if skew and it == 2:
                assert min(len(parts[0]), len(parts[1])) == 0
        fAx, fl, fu = r0[('single', it)]
        A_full = sp.csc_matrix((fAx, findices, findptr), shape=(fr, fc))
        As, ls, us_ = _submatrix(A_full, fl, fu, idx, M, R, nu, nfin)
        assert (n_rows, n_cols) == As.shape and np.array_equal(indptr, As.indptr) and np.array_equal(indices, As.indices)
        fin = indices < nfin
        assert np.array_equal(Ax[~fin], As.data[~fin])            # bitwise: same kernel, same samples
        assert np.allclose(Ax[fin], As.data[fin], rtol=1e-12, atol=1e-14)
        assert np.array_equal(l[nfin:], ls[nfin:]) and np.array_equal(u[nfin:], us_[nfin:])
        assert np.allclose(l[:nfin], ls[:nfin], rtol=1e-12, atol=1e-14) and np.allclose(u[:nfin], us_[:nfin], rtol=1e-12, atol=1e-14)


def _qp_worker(rank, world, initfile, M, out):
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist
    from riskaversetrajopt_b200 import _lib, dist as sd
    from riskaversetrajopt_b200.device_path import DevicePath
    from riskaversetrajopt_b200.device_qp import DeviceQP
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{initfile}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)

    def make(first, cnt):
        p = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, cnt, M_global=M, sample_offset=first, device=rank)
        p.set_params_drone(dp, dp.OSQP_TOL)
        p.set_samples_drone(masses[first:first + cnt], DWs[first:first + cnt], obs_Qs[first:first + cnt])
        return p
    n = 62 + M
    P = sp.lil_matrix((n, n))
    P[:60, :60] = np.kron(np.eye(20), 2 * dp.dt * np.asarray(dp.R))
    P[n - 2, n - 2] = 1e4
    P = sp.csc_matrix(P)
    q = np.zeros(n); q[-2] = 1e4
    first, cnt = sd.shard_range(M, world, rank)
    path = make(first, cnt)
    asm = sd.ShardedAssembler(path, mode='sharded')          # compact row block per rank, global means
    us = np.tile(np.array([0.01, 0.01, 0.0]), (20, 1))
    dq = DeviceQP(path, group=dist.group.WORLD, eps_abs=1e-4, eps_rel=1e-4)
    dq.setup(P, q, asm.step(us, 2))
    single = sq = None
    if rank == 0:
        single = make(0, M)
        sq = DeviceQP(single, eps_abs=1e-4, eps_rel=1e-4).setup(P, q, single.assemble(us, 2))
    res = []
    for it in range(4):
        dq.update(asm.step(us, it))
        r = dq.solve()
        rec = dict(u=r.u, t=r.t, slack=r.slack, iters=r.info.iter, status=r.info.status)
        if rank == 0:
            sq.update(single.assemble(us, it))
            r1 = sq.solve()
            rec.update(u1=r1.u, t1=r1.t, iters1=r1.info.iter, status1=r1.info.status)
        res.append(rec)
        us = np.reshape(r.u, (3, 20), 'F').T
    out[rank] = res
    dist.destroy_process_group()


def test_two_ranks_solve_the_qp_on_sharded_rows():
    """The device ADMM over row blocks that stay sharded (one (nu + 4)-double all-reduce per iteration): every
    rank gets the solution one GPU computes over all samples."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    mgr = mp.Manager()
    out = mgr.dict()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_qp_worker, args=(2, os.path.join(d, "init"), 301, out), nprocs=2, join=True)
    for it, (a, b) in enumerate(zip(out[0], out[1])):
        assert a['status'] == b['status'] == a['status1'] == 'solved'
        assert a['iters'] == b['iters'] and np.array_equal(a['u'], b['u']) and a['t'] == b['t']   # replicated dense state
        assert abs(a['iters'] - a['iters1']) <= 10
        assert np.max(np.abs(a['u'] - a['u1'])) < (1e-7 if it == 0 else 1e-4) and abs(a['t'] - a['t1']) < 1e-4
