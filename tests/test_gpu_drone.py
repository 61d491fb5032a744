"""Parity of the CUDA path (through the C ABI) with the oracles -- drone.

Tolerances (BASELINE.json north_star): FP64 <= 1e-9 relative on values, pattern
and indices bit-exact; FP32 <= 1e-4.
"""
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
RTOL64, RTOL32 = 1e-9, 1e-4


def _models(seed0, M, method='saa', alpha=0.1, variant='risk', precision='fp64'):
    from oracle.oracle_b import DroneOracleB
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_risk import Model
    DWs, masses, obs_Qs = (x[:M] for x in seed0)
    return (Model(dp.S, DWs, masses, obs_Qs, method, alpha, variant=variant, precision=precision),
            DroneOracleB(dp.S, DWs, masses, obs_Qs, method, alpha, variant))


def _big_samples(M, seed=0):
    """Config 4 synthetic samples in the reference's distributions (vectorised)."""
    from riskaversetrajopt_b200.drone import drone_params as dp
    rs = np.random.RandomState(seed)
    masses = rs.uniform(dp.mass_nom - dp.mass_delta, dp.mass_nom + dp.mass_delta, M)
    obs_Qs = np.zeros((M, 3, 3, 3))
    for o in range(3):
        for d in range(3):
            obs_Qs[:, o, d, d] = 1. / (dp.obs_radii[o] + rs.uniform(-dp.obs_radii_deltas, dp.obs_radii_deltas, M))**2
    DWs = np.sqrt(dp.dt) * rs.randn(M, dp.S, 6)
    return DWs, masses, obs_Qs


def _check(A, l, u, Ar, lr, ur, rtol):
    assert A.shape == Ar.shape
    assert A.indptr.dtype == Ar.indptr.dtype and A.indices.dtype == Ar.indices.dtype
    assert np.array_equal(A.indptr, Ar.indptr) and np.array_equal(A.indices, Ar.indices)
    assert rel_err(A.data, Ar.data, 1e-12 if rtol < 1e-6 else 1e-6) < rtol
    assert np.array_equal(np.isinf(l), np.isinf(lr)) and np.array_equal(np.isinf(u), np.isinf(ur))
    f = np.isfinite(lr)
    assert np.allclose(l[f], lr[f], rtol=rtol, atol=rtol * 1e-2)
    assert np.allclose(u, ur, rtol=rtol, atol=rtol * 1e-2)


@pytest.mark.parametrize("method,variant", [("saa", "risk"), ("saa", "times"),
                                            ("baseline", "risk"), ("baseline", "times")])
@pytest.mark.parametrize("scp_iter", [0, 1, 2, 7])
def test_reference_config_matches_oracle(drone_seed0, method, variant, scp_iter):
    model, ref = _models(drone_seed0, 50, method, 0.1, variant)
    us = model.initial_guess_us_mat() + 0.2 * np.random.RandomState(scp_iter).randn(20, 3)
    _check(*model.get_constraints_coeffs(us, scp_iter), *ref.get_constraints_coeffs(us, scp_iter), RTOL64)


@pytest.mark.parametrize("src", ["", "ref_"], ids=["oracle_a", "reference_exec"])
def test_golden_m50(drone_seed0, src):
    g = np.load(os.path.join(G, src + "drone_M50_saa_iter2.npz"))
    model, _ = _models(drone_seed0, 50)
    assert np.array_equal(model.initial_guess_us_mat(), g["us"])
    A, l, u = model.get_constraints_coeffs(g["us"], 2)
    assert A.shape == tuple(g["shape"])
    assert np.array_equal(A.indptr, g["indptr"]) and np.array_equal(A.indices, g["indices"])
    assert rel_err(A.data, g["data"]) < RTOL64
    assert np.allclose(u, g["u"], rtol=RTOL64, atol=1e-12)
    assert np.allclose(l[np.isfinite(l)], g["l"][np.isfinite(g["l"])], rtol=RTOL64, atol=1e-12)
    Xs = model.us_to_state_trajectories(g["us"])
    assert Xs.shape == (50, 21, 6)
    assert np.allclose(Xs[:3], g["Xs_first3"], rtol=1e-11, atol=1e-13)
    sat, Z = model.monte_carlo_constraints(g["us"])
    assert np.allclose(Z, g["Z"], rtol=RTOL64, atol=1e-12)
    assert np.array_equal(sat, g["Z"] <= 1e-6)


@pytest.mark.parametrize("src", ["", "ref_"], ids=["oracle_a", "reference_exec"])
def test_golden_m8_branches(drone_seed0, src):
    g = np.load(os.path.join(G, src + "drone_M8_branches.npz"))
    for method in ('saa', 'baseline'):
        for variant in ('risk', 'times'):
            model, _ = _models(drone_seed0, 8, method, 0.05, variant)
            for it in (0, 2):
                A, l, u = model.get_constraints_coeffs(g["us"], it)
                k = f"{method}_{variant}_{it}"
                assert np.array_equal(A.indptr, g[k + "_indptr"])
                assert np.array_equal(A.indices, g[k + "_indices"])
                assert rel_err(A.data, g[k + "_data"]) < RTOL64
                assert np.allclose(u, g[k + "_u"], rtol=RTOL64, atol=1e-12)
                f = np.isfinite(g[k + "_l"])
                assert np.array_equal(np.isfinite(l), f)
                assert np.allclose(l[f], g[k + "_l"][f], rtol=RTOL64, atol=1e-12)


@pytest.mark.parametrize("M", [1, 2, 15, 16, 17, 31, 33, 100, 1000])
def test_ragged_sample_counts(M):
    """Tile size is 16 samples per warp: cover partial tiles and single samples."""
    from oracle.oracle_b import DroneOracleB
    from riskaversetrajopt_b200.drone.drone_risk import Model
    DWs, masses, obs_Qs = _big_samples(M, seed=M)
    model = Model(20, DWs, masses, obs_Qs, 'saa', 0.2)
    ref = DroneOracleB(20, DWs, masses, obs_Qs, 'saa', 0.2)
    us = np.random.RandomState(1).randn(20, 3)
    _check(*model.get_constraints_coeffs(us, 3), *ref.get_constraints_coeffs(us, 3), RTOL64)
    Xs = model.us_to_state_trajectories(us)
    assert np.allclose(Xs, ref.rollout(us)[0], rtol=1e-11, atol=1e-13)


def test_zero_controls_and_zero_velocity_start(drone_seed0):
    """|v| = 0 at k = 0 (d|v|/dv = 0 there) and u = 0: structural zeros must be kept
    as explicit entries, values equal the oracle's."""
    model, ref = _models(drone_seed0, 20)
    us = np.zeros((20, 3))
    _check(*model.get_constraints_coeffs(us, 2), *ref.get_constraints_coeffs(us, 2), RTOL64)


def test_repeated_calls_are_bitwise_deterministic(drone_seed0):
    model, _ = _models(drone_seed0, 50)
    us = np.random.RandomState(3).randn(20, 3)
    A1, l1, u1 = model.get_constraints_coeffs(us, 2)
    A2, l2, u2 = model.get_constraints_coeffs(us, 2)
    assert np.array_equal(A1.data, A2.data) and np.array_equal(u1, u2) and np.array_equal(l1, l2)


def test_iteration_switch_rewrites_constants(drone_seed0):
    """relaxed (scp_iter < 2) -> normal -> relaxed on the same buffers."""
    model, ref = _models(drone_seed0, 10)
    us = np.random.RandomState(4).randn(20, 3)
    for it in (0, 2, 1, 5):
        _check(*model.get_constraints_coeffs(us, it), *ref.get_constraints_coeffs(us, it), RTOL64)


def test_dense_rows_and_objective(drone_seed0):
    from oracle.oracle_a import DroneOracleA
    from riskaversetrajopt_b200.drone import drone_params as dp
    model, _ = _models(drone_seed0, 6)
    a = DroneOracleA(dp.S, *(x[:6] for x in drone_seed0), 'saa', 0.1)
    us = model.initial_guess_us_mat()
    D, lo, up = model.get_all_constraints_coeffs_all(us)
    Da, loa, upa = a.get_all_constraints_coeffs_all(us)
    assert D.shape == Da.shape and np.allclose(D, Da, rtol=RTOL64, atol=1e-13)
    assert np.allclose(up, upa, rtol=RTOL64, atol=1e-13)
    P, q = model.get_objective_coeffs()
    Pa, qa = a.get_objective_coeffs()
    assert (P != Pa).nnz == 0 and np.array_equal(q, qa)
    assert np.array_equal(P.indices, Pa.indices) and np.array_equal(P.indptr, Pa.indptr)


def _per_entry_fp32(A, Ar):
    """north_star: FP32 within 1e-4 RELATIVE -- per entry, on every entry above 1e-6 of its
    column's scale (not relative to the matrix norm)."""
    for c in range(A.shape[1]):
        lo, hi = Ar.indptr[c], Ar.indptr[c + 1]
        if hi == lo:
            continue
        ref, got = Ar.data[lo:hi], A.data[lo:hi]
        big = np.abs(ref) > 1e-6 * np.max(np.abs(ref))
        assert np.max(np.abs(got[big] - ref[big]) / np.abs(ref[big])) < RTOL32, c
        assert np.max(np.abs(got - ref)) <= 1e-6 * np.max(np.abs(ref)) + 1e-300, c


def test_fp32_mode_within_1e4(drone_seed0):
    model, ref = _models(drone_seed0, 50, precision='fp32')
    us = model.initial_guess_us_mat() + 0.2 * np.random.RandomState(0).randn(20, 3)
    A, l, u = model.get_constraints_coeffs(us, 2)
    Ar, lr, ur = ref.get_constraints_coeffs(us, 2)
    assert np.array_equal(A.indices, Ar.indices) and np.array_equal(A.indptr, Ar.indptr)
    _per_entry_fp32(A, Ar)
    f = np.isfinite(ur)
    assert np.array_equal(np.isfinite(u), f)
    assert np.max(np.abs(u[f] - ur[f]) / np.maximum(np.abs(ur[f]), 1e-30)) < RTOL32
    fl = np.isfinite(lr)
    assert np.max(np.abs(l[fl] - lr[fl]) / np.maximum(np.abs(lr[fl]), 1e-30)) < RTOL32
    # default (copy=True): doubles like the reference's boundary; copy=False: the float32 staging buffers as they are
    assert A.data.dtype == np.float64 and u.dtype == np.float64
    A32, l32, u32 = model.get_constraints_coeffs(us, 2, copy=False)
    assert A32.data.dtype == np.float32 and u32.dtype == np.float32 and l32.dtype == np.float32
    assert np.array_equal(A32.data.astype(np.float64), A.data) and np.array_equal(A32.indices, A.indices)
    assert np.array_equal(u32.astype(np.float64), u)


def test_large_M_sampled_parity_and_properties():
    """M = 2*10^5 (config-4 style): sampled sub-runs against the oracle, plus
    size-independent properties."""
    import torch
    from oracle.oracle_b import DroneOracleB
    from riskaversetrajopt_b200.drone.drone_risk import Model
    M = 200_000
    DWs, masses, obs_Qs = _big_samples(M, seed=7)
    model = Model(20, DWs, masses, obs_Qs, 'saa', 0.1)
    us = model.initial_guess_us_mat() + 0.1 * np.random.RandomState(2).randn(20, 3)
    b = model.path.assemble(us, 2)
    Ax, l, u = (b[k].cpu().numpy() for k in ('Ax', 'l', 'u'))
    n_rows, n_cols, indptr, indices = model.path.pattern()
    assert Ax.size == 1263 * M + 180 == indptr[-1]
    # sampled samples: their sub-runs in every u column + their bounds
    idx = np.unique(np.concatenate([[0, 1, 15, 16, M - 1, M - 17], np.random.RandomState(0).randint(0, M, 40)]))
    ref = DroneOracleB(20, DWs[idx], masses[idx], obs_Qs[idx], 'saa', 0.1)
    _, _, _, g_du, g_up = ref.per_sample(us)
    row_s0 = 6 + 1 + M
    for n, i in enumerate(idx):
        assert rel_err(u[row_s0 + i * 60: row_s0 + (i + 1) * 60], 0.01 * g_up[n].reshape(-1)) < RTOL64
        for j in (0, 5, 18):
            for a in (0, 1):
                c = j * 3 + a
                L = 19 - j
                nfin = 2
                start = indptr[c] + nfin + i * 3 * L
                got = Ax[start:start + 3 * L].reshape(3, L)
                want = 0.01 * g_du[n, :, j + 1:, c]
                assert rel_err(got, want) < RTOL64
                assert np.array_equal(indices[start:start + 3 * L].reshape(3, L)[0],
                                      row_s0 + i * 60 + np.arange(j + 1, 20))
    # property: the mean rows are the mean over ALL samples (chunked oracle)
    fin = np.zeros(6)
    for s in range(0, M, 50_000):
        r = DroneOracleB(20, DWs[s:s + 50_000], masses[s:s + 50_000], obs_Qs[s:s + 50_000])
        Xs, _ = r.rollout(us)
        fin += (Xs[:, 20, :] - 0).sum(axis=0)
    # l[0:6] = mean(-(x_S - x_f) + J u); check the affine identity via a second iterate:
    b2 = model.path.assemble(np.zeros((20, 3)), 2)
    l0 = b2['l'][:6].cpu().numpy()            # at u = 0: l = -mean(x_S(0) - x_f)
    fin0 = np.zeros(6)
    for s in range(0, M, 50_000):
        r = DroneOracleB(20, DWs[s:s + 50_000], masses[s:s + 50_000], obs_Qs[s:s + 50_000])
        fin0 += r.rollout(np.zeros((20, 3)))[0][:, 20, :].sum(axis=0)
    assert np.allclose(l0, -fin0 / M, rtol=1e-9, atol=1e-12)
    # property: constants
    assert np.all(np.isneginf(l[6:row_s0 + 60 * M + 1]))
    assert Ax[indptr[60 + M + 1]] == M * 0.1          # t column, CVaR row
    assert np.all(Ax[indptr[60 + M + 1] + 1: indptr[60 + M + 2]] == -0.01)
    # property: K5 terms agree with Z written by the assemble kernel
    Z = torch.empty(M, dtype=torch.float64, device='cuda')
    model.path.assemble(us, 2, Z=Z)
    Zc, out3 = model.path.cvar_terms(us, t_risk=-0.5, sat_tol=1e-6)
    from riskaversetrajopt_b200.drone import drone_params as dp
    assert torch.allclose(Z - dp.OSQP_TOL, Zc, rtol=1e-12, atol=1e-13)
    Zh = Zc.cpu().numpy()
    assert np.isclose(out3[0].item(), np.maximum(Zh + 0.5, 0).sum(), rtol=1e-10)
    assert out3[1].item() == np.count_nonzero(Zh <= 1e-6) and out3[2].item() == Zh.max()


def test_monte_carlo_statistics_on_device(drone_seed0):
    """V@R by selection and AV@R in closed form == the reference's sort / LP definitions."""
    import torch
    from riskaversetrajopt_b200 import montecarlo as mc
    from riskaversetrajopt_b200.drone.drone_risk import Model
    M = 20_000
    DWs, masses, obs_Qs = _big_samples(M, seed=3)
    model = Model(20, DWs, masses, obs_Qs, 'saa', 0.1)
    us = model.initial_guess_us_mat() + 0.3 * np.random.RandomState(8).randn(20, 3)
    Z, _ = model.path.cvar_terms(us, 0.0, 1e-6)
    Zh = Z.cpu().numpy()
    for alpha in (0.05, 0.1, 0.3):
        xth = int(np.floor(alpha * M))
        assert mc.monte_carlo_var(model.path, Z, alpha) == np.sort(Zh)[M - xth - 1]        # drone_main_plot.py:649-651
        # the LP of drone_risk.py:664-695 minimises t + mean((Z-t)^+)/alpha over t: brute force on a grid of
        # order statistics
        cand = np.sort(Zh)[M - xth - 50:M - xth + 50]
        obj = np.array([t + np.maximum(Zh - t, 0).mean() / alpha for t in cand])
        assert np.isclose(mc.monte_carlo_avar(model.path, Z, alpha), obj.min(), rtol=1e-9, atol=1e-12)
        assert np.isclose(mc.monte_carlo_avar(model.path, Z, alpha, t_risk=0.1),
                          0.1 + np.maximum(Zh - 0.1, 0).mean() / alpha, rtol=1e-12)
    assert mc.fraction_satisfied(Z) == np.mean(Zh <= 1e-6)
    assert isinstance(Z, torch.Tensor) and Z.is_cuda


@pytest.mark.parametrize("M,scp_iter", [(50, 2), (37, 0), (1000, 5)])
def test_factored_record_expands_to_identical_entries(drone_seed0, M, scp_iter):
    """FACTOR + EXPAND (the multi-GPU gather path) on one GPU: the expanded u-column block is
    bitwise the one the FULL kernel writes; bounds and mean sums agree too."""
    import torch
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_risk import Model
    DWs, masses, obs_Qs = _big_samples(M, seed=11) if M > 50 else tuple(x[:M] for x in drone_seed0)
    model = Model(20, DWs, masses, obs_Qs, 'saa', 0.1)
    p = model.path
    us = model.initial_guess_us_mat() + 0.2 * np.random.RandomState(M).randn(20, 3)
    full = {k: v.clone() for k, v in p.assemble(us, scp_iter).items() if torch.is_tensor(v)}
    sums_full = p.mean_sums.clone()
    n_sp, n_p = p.factored_sizes()
    assert n_sp == 380 * M and n_p == 46 * M
    fsp = torch.full((n_sp,), float('nan'), dtype=torch.float64, device='cuda')
    fp = torch.full((n_p,), float('nan'), dtype=torch.float64, device='cuda')
    n_rows, _, nnz = p.pattern_sizes()
    Ax = torch.zeros(nnz, dtype=torch.float64, device='cuda')
    u = torch.full((n_rows,), float('nan'), dtype=torch.float64, device='cuda')
    p.linearize_factored(us, scp_iter, fsp, fp, u)
    assert not torch.isnan(fsp).any() and not torch.isnan(fp).any()
    # expand in two pieces with an odd split point, as a matrix owner does for several ranks
    cut = M // 3
    p.expand_factored(scp_iter, fsp, fp, 0, cut, Ax)
    p.expand_factored(scp_iter, fsp, fp, cut, M - cut, Ax)
    torch.cuda.synchronize()
    n_rows_, n_cols, indptr, indices = p.pattern()
    for c in range(60):
        j, a = divmod(c, 3)
        if a == 2 or j > 18:
            continue
        lo, hi = indptr[c] + 2, indptr[c + 1] - 1             # skip the final rows and the control row
        assert torch.equal(Ax[lo:hi], full['Ax'][lo:hi]), c
    if scp_iter >= 2:
        lo = 7 + M
        assert torch.equal(u[lo:lo + 60 * M], full['u'][lo:lo + 60 * M])
    # same samples, same per-sample values; the reduction tree may differ between the kernels
    assert torch.allclose(p.mean_sums, sums_full, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("S", [3, 5, 12, 30, 32])
@pytest.mark.parametrize("method", ["saa", "baseline"])
def test_other_horizons_run_the_generic_kernels(S, method):
    """The reference's Model takes S as a constructor argument (drone/drone_risk.py:70-83, dt = T / S).
    S = 20 runs the tuned kernels, every other horizon the generic ones (csrc/generic_kernels.cuh):
    same outputs against the oracle, including relaxed iterations, rollout and the CVaR terms."""
    from oracle.oracle_b import DroneOracleB
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_risk import Model
    M = 37
    rs = np.random.RandomState(S)
    masses = rs.uniform(dp.mass_nom - dp.mass_delta, dp.mass_nom + dp.mass_delta, M)
    obs_Qs = np.zeros((M, 3, 3, 3))
    for o in range(3):
        for d in range(3):
            obs_Qs[:, o, d, d] = 1. / (dp.obs_radii[o] + rs.uniform(-dp.obs_radii_deltas, dp.obs_radii_deltas, M))**2
    DWs = np.sqrt(dp.T / S) * rs.randn(M, S, 6)
    model = Model(S, DWs, masses, obs_Qs, method, 0.2)
    ref = DroneOracleB(S, DWs, masses, obs_Qs, method, 0.2)
    us = model.initial_guess_us_mat() + 0.3 * rs.randn(S, 3)
    for it in (0, 2, 5):
        _check(*model.get_constraints_coeffs(us, it), *ref.get_constraints_coeffs(us, it), RTOL64)
    assert np.allclose(model.us_to_state_trajectories(us), ref.rollout(us)[0], rtol=1e-11, atol=1e-13)
    sat, Z = model.monte_carlo_constraints(us)
    assert np.allclose(Z, ref.monte_carlo_constraints(us)[1], rtol=1e-10, atol=1e-12)
    avar = model.monte_carlo_avar(us, 0.1)
    assert np.isclose(avar, 0.1 + np.mean(np.maximum(Z - 0.1, 0)) / 0.2, rtol=1e-10)


def test_generic_horizon_vs_reference_execution():
    """S = 12, M = 6 through the reference's own code (tests/golden/ref_drone_S12.npz)."""
    from riskaversetrajopt_b200.drone.drone_risk import Model
    g = np.load(os.path.join(G, "ref_drone_S12.npz"))
    model = Model(12, g["DWs"], g["masses"], g["obs_Qs"], 'saa', 0.1)
    for it in (0, 2):
        A, l, u = model.get_constraints_coeffs(g["us"], it)
        k = f"iter{it}"
        assert np.array_equal(A.indptr, g[k + "_indptr"]) and np.array_equal(A.indices, g[k + "_indices"])
        assert rel_err(A.data, g[k + "_data"]) < RTOL64
        assert np.allclose(u, g[k + "_u"], rtol=RTOL64, atol=1e-12)
        f = np.isfinite(g[k + "_l"])
        assert np.allclose(l[f], g[k + "_l"][f], rtol=RTOL64, atol=1e-12)
