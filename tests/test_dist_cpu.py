"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes, the C oracle standing in
for the kernel (the CUDA path itself is covered by tests/test_gpu_multi.py)."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from riskaversetrajopt_b200.dist import shard_range


def test_shard_range_partitions_everything():
    for M, W in ((50, 2), (50, 4), (50, 8), (10**6, 8), (17, 3), (8, 8)):
        blocks = [shard_range(M, W, r) for r in range(W)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == M
        for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
            assert f0 + c0 == f1
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        shard_range(3, 4, 0)


def _worker(rank, world, initfile, M, out):
    from oracle import cpu_port
    from riskaversetrajopt_b200 import dist as sd
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)
    us0 = np.random.RandomState(1).randn(20, 3) if rank == 0 else np.zeros((20, 3))
    us = sd.broadcast_controls(us0, 0)                       # rank 0's iterate everywhere
    first, cnt = sd.shard_range(M, world, rank)
    sl = slice(first, first + cnt)
    Ax, ub, sums, Z, off = cpu_port.drone_assemble(us, masses[sl], DWs[sl], obs_Qs[sl])
    t = torch.from_numpy(sums.copy())
    sd.all_reduce_sums(t)                                     # what NCCL does on the GPUs
    out[rank] = dict(us=us, first=first, cnt=cnt, Ax=Ax, ub=ub, sums=t.numpy(), off=off)
    dist.destroy_process_group()


def test_two_rank_shards_reassemble_to_the_global_problem(built_lib):
    from oracle import cpu_port
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    from riskaversetrajopt_b200.pattern import csc_pattern
    M, W = 21, 2
    mgr = mp.Manager()
    out = mgr.dict()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(W, os.path.join(d, "init"), M, out), nprocs=W, join=True)
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)
    us = out[0]["us"]
    assert np.array_equal(us, out[1]["us"])
    gAx, gub, gsums, _, goff = cpu_port.drone_assemble(us, masses, DWs, obs_Qs)
    # mean-row sums: all-reduce of the shard sums == global sums (up to summation order)
    assert np.allclose(out[0]["sums"], gsums, rtol=1e-13, atol=1e-13)
    assert np.array_equal(out[0]["sums"], out[1]["sums"])
    # row blocks: rank r's sub-run of column (j,a) sits at start + first_r * len_c
    n_rows, n_cols, indptr, indices = csc_pattern('drone', 'saa', 20, M)
    merged = np.zeros_like(gAx)
    for r in range(W):
        o = out[r]
        for a in range(2):
            for j in range(19):
                L = 3 * (19 - j)
                src = o["off"][a * 19 + j]
                dst = goff[a * 19 + j] + o["first"] * L
                assert goff[a * 19 + j] == indptr[j * 3 + a] + 2
                merged[dst:dst + o["cnt"] * L] = o["Ax"][src:src + o["cnt"] * L]
        assert np.array_equal(gub[o["first"] * 60:(o["first"] + o["cnt"]) * 60], o["ub"])
    assert np.array_equal(merged, gAx)


def test_tie_quotas_of_the_exact_global_selection():
    """Places left after the values above the threshold go to the ties in rank order."""
    from riskaversetrajopt_b200.dist import tie_quotas
    assert tie_quotas([(3, 0), (2, 0)], 5) == [(3, 0), (2, 0)]
    assert tie_quotas([(3, 4), (2, 5), (0, 1)], 8) == [(6, 3), (2, 0), (0, 0)]          # 3 places, all to rank 0
    assert tie_quotas([(0, 2), (0, 2), (0, 2)], 5) == [(2, 2), (2, 2), (1, 1)]          # everything ties
    assert tie_quotas([(0, 0), (7, 1)], 8) == [(0, 0), (8, 1)]                          # a rank that contributes nothing
    rs = np.random.RandomState(0)
    for _ in range(50):
        z = np.round(rs.randn(200) * 2)
        cuts = np.sort(rs.choice(np.arange(1, 200), 3, replace=False))
        parts = np.split(z, cuts)
        K = int(rs.randint(1, 201))
        thr = np.sort(z)[::-1][K - 1]
        q = tie_quotas([((p > thr).sum(), (p == thr).sum()) for p in parts], K)
        assert sum(c for c, _ in q) == K
        # the union of the per-rank picks is the global top K with ties to the smaller global index
        order = np.lexsort((np.arange(200), -z))[:K]
        got, off = [], 0
        for p, (c, take) in zip(parts, q):
            loc = np.concatenate([np.flatnonzero(p > thr), np.flatnonzero(p == thr)[:take]])
            got.append(np.sort(loc) + off); off += len(p)
        assert np.array_equal(np.concatenate(got), np.sort(order))
    import pytest
    with pytest.raises(ValueError):
        tie_quotas([(3, 0), (3, 0)], 5)
