"""Parity of the CUDA path (through the C ABI) with the oracles -- car + pedestrian."""
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
RTOL64, RTOL32 = 1e-9, 1e-4


def _seed0(M=50, method='saa'):
    from riskaversetrajopt_b200.car.driving import sample_uncertain_parameters
    state = np.random.get_state()
    np.random.seed(0)
    out = sample_uncertain_parameters(M, method)
    np.random.set_state(state)
    return out


def _check(A, l, u, Ar, lr, ur, rtol=RTOL64):
    assert A.shape == Ar.shape
    assert A.indptr.dtype == Ar.indptr.dtype and A.indices.dtype == Ar.indices.dtype
    assert np.array_equal(A.indptr, Ar.indptr) and np.array_equal(A.indices, Ar.indices)
    assert rel_err(A.data, Ar.data, 1e-12) < rtol
    assert np.array_equal(np.isnan(l), np.isnan(lr)) and np.array_equal(np.isinf(l), np.isinf(lr))
    f = np.isfinite(lr)
    assert np.allclose(l[f], lr[f], rtol=rtol, atol=rtol * 1e-2)
    assert np.allclose(u, ur, rtol=rtol, atol=rtol * 1e-2)


@pytest.mark.parametrize("scp_iter", [0, 1, 2, 9])
def test_reference_config_matches_oracle(scp_iter):
    from oracle.oracle_b import CarOracleB
    from riskaversetrajopt_b200.car.driving import Model
    s = _seed0()
    model, ref = Model(50, 'saa', 0.05, samples=s), CarOracleB(*s, 'saa', 0.05)
    us = model.initial_guess_us_mat() + 0.3 * np.random.RandomState(scp_iter).randn(20, 2)
    _check(*model.get_constraints_coeffs(us, scp_iter), *ref.get_constraints_coeffs(us, scp_iter))


def test_model_constructor_draws_like_the_reference():
    from riskaversetrajopt_b200.car.driving import Model
    np.random.seed(0)
    m = Model(50, 'saa', 0.05)
    s = _seed0()
    assert all(np.array_equal(a, b) for a, b in zip(
        (m.states_init, m.omegas_speed, m.omegas_repulsive, m.DWs), s))


@pytest.mark.parametrize("src", ["", "ref_"], ids=["oracle_a", "reference_exec"])
def test_golden(src):
    from riskaversetrajopt_b200.car.driving import Model
    g = np.load(os.path.join(G, src + "car_M50_saa.npz"))
    model = Model(50, 'saa', 0.05, samples=_seed0())
    assert np.array_equal(model.initial_guess_us_mat(), g["us0"])
    for name, us, it in (("iter0", g["us0"], 0), ("iter1", g["us1"], 1), ("iter2", g["us1"], 2)):
        A, l, u = model.get_constraints_coeffs(us, it)
        assert A.shape == tuple(g[name + "_shape"])
        assert np.array_equal(A.indptr, g[name + "_indptr"]) and np.array_equal(A.indices, g[name + "_indices"])
        assert rel_err(A.data, g[name + "_data"]) < RTOL64
        assert np.allclose(u, g[name + "_u"], rtol=RTOL64, atol=1e-11)
        gl = g[name + "_l"]
        assert np.array_equal(np.isnan(l), np.isnan(gl)) and np.array_equal(np.isinf(l), np.isinf(gl))
        f = np.isfinite(gl)
        assert np.allclose(l[f], gl[f], rtol=RTOL64, atol=1e-11)
    Xs = model.us_to_state_trajectories(g["us1"])
    assert Xs.shape == (50, 21, 8)
    assert np.allclose(Xs[:3], g["Xs_first3"], rtol=1e-11, atol=1e-12)
    sat, Z = model.monte_carlo_constraints(g["us1"])
    assert np.allclose(Z, g["Z"], rtol=RTOL64, atol=1e-11)


def test_baseline_method():
    from oracle.oracle_b import CarOracleB
    from riskaversetrajopt_b200.car.driving import Model
    s = _seed0(20, 'baseline')
    model, ref = Model(20, 'baseline', 0.05, samples=s), CarOracleB(*s, 'baseline', 0.05)
    us = model.initial_guess_us_mat() + 0.2 * np.random.RandomState(1).randn(20, 2)
    _check(*model.get_constraints_coeffs(us, 1), *ref.get_constraints_coeffs(us, 1))


@pytest.mark.parametrize("M", [3, 15, 16, 17, 33, 200])
def test_ragged_sample_counts(M):
    from oracle.oracle_b import CarOracleB
    from riskaversetrajopt_b200.car.driving import Model
    s = tuple(x[:M] for x in _seed0(200))
    model, ref = Model(M, 'saa', 0.1, samples=s), CarOracleB(*s, 'saa', 0.1)
    us = model.initial_guess_us_mat() + np.random.RandomState(M).randn(20, 2)
    for it in (0, 1):
        _check(*model.get_constraints_coeffs(us, it), *ref.get_constraints_coeffs(us, it))
    assert np.allclose(model.us_to_state_trajectories(us), ref.rollout(us), rtol=1e-11, atol=1e-12)


def test_iteration_order_0_1_2_and_back():
    from oracle.oracle_b import CarOracleB
    from riskaversetrajopt_b200.car.driving import Model
    s = tuple(x[:10] for x in _seed0())
    model, ref = Model(10, 'saa', 0.05, samples=s), CarOracleB(*s, 'saa', 0.05)
    us = model.initial_guess_us_mat()
    for it in (0, 1, 2, 0, 3):
        _check(*model.get_constraints_coeffs(us, it), *ref.get_constraints_coeffs(us, it))


def test_ego_state_must_be_shared():
    from riskaversetrajopt_b200._lib import SaaError
    from riskaversetrajopt_b200.car.driving import Model
    s = list(_seed0(8))
    s[0] = s[0].copy(); s[0][3, 0] += 1.0
    with pytest.raises(SaaError):
        Model(8, 'saa', 0.05, samples=tuple(s))


def test_fp32_mode_within_1e4():
    from oracle.oracle_b import CarOracleB
    from riskaversetrajopt_b200.car.driving import Model
    s = _seed0()
    model, ref = Model(50, 'saa', 0.05, samples=s, precision='fp32'), CarOracleB(*s, 'saa', 0.05)
    us = model.initial_guess_us_mat() + 0.2 * np.random.RandomState(0).randn(20, 2)
    A, l, u = model.get_constraints_coeffs(us, 2)
    Ar, lr, ur = ref.get_constraints_coeffs(us, 2)
    assert np.array_equal(A.indices, Ar.indices)
    # north_star: 1e-4 relative PER ENTRY, on every entry above 1e-6 of its column's scale
    for c in range(A.shape[1]):
        lo, hi = Ar.indptr[c], Ar.indptr[c + 1]
        ref_c, got = Ar.data[lo:hi], A.data[lo:hi]
        big = np.abs(ref_c) > 1e-6 * np.max(np.abs(ref_c))
        assert np.max(np.abs(got[big] - ref_c[big]) / np.abs(ref_c[big])) < RTOL32, c
    f = np.isfinite(ur)
    assert np.max(np.abs(u[f] - ur[f]) / np.maximum(np.abs(ur[f]), 1e-30)) < RTOL32


def test_large_M_sampled_parity():
    import torch
    from oracle.oracle_b import CarOracleB
    from riskaversetrajopt_b200.car.driving import Model, sample_uncertain_parameters
    M = 100_000
    st = np.random.get_state(); np.random.seed(5)
    s = sample_uncertain_parameters(M, 'saa'); np.random.set_state(st)
    model = Model(M, 'saa', 0.05, samples=s)
    us = model.initial_guess_us_mat() + 0.1 * np.random.RandomState(2).randn(20, 2)
    b = model.path.assemble(us, 2)
    Ax, u = b['Ax'].cpu().numpy(), b['u'].cpu().numpy()
    n_rows, n_cols, indptr, indices = model.path.pattern()
    assert Ax.size == 423 * M + 159
    idx = np.unique(np.concatenate([[0, 15, 16, M - 1], np.random.RandomState(0).randint(0, M, 30)]))
    ref = CarOracleB(*(x[idx] for x in s), 'saa', 0.05)
    _, _, _, g_du, g_up, g = ref.per_sample(us)
    row_s0 = 4 + 1 + M
    for n, i in enumerate(idx):
        assert rel_err(u[row_s0 + i * 20: row_s0 + (i + 1) * 20], g_up[n]) < RTOL64
        for j in (0, 7, 18):
            for c in (0, 1):
                col, L = j * 2 + c, 19 - j
                start = indptr[col] + 3 + i * L
                assert rel_err(Ax[start:start + L], g_du[n, j + 1:, col]) < RTOL64
                assert np.array_equal(indices[start:start + L], row_s0 + i * 20 + np.arange(j + 1, 20))
    Zc, out3 = model.path.cvar_terms(us, t_risk=-1.0)
    Zh = Zc.cpu().numpy()
    assert np.allclose(Zh[idx], g.max(axis=1) - 3e-4, rtol=1e-10, atol=1e-12)
    assert np.isclose(out3[0].item(), np.maximum(Zh + 1.0, 0).sum(), rtol=1e-10)


@pytest.mark.parametrize("M,method", [(50, 'baseline'), (8, 'baseline'), (1, 'saa'), (2, 'saa'), (3, 'saa'),
                                      (1, 'baseline')])
def test_relaxation_edge_cases_vs_reference_execution(M, method):
    """car/driving.py:411-415 at scp_iter 0: rows < n_x = 8 survive.  With the baseline (what the
    reference's driver runs first, :536-537) those are sample 0's first four separation rows WITH
    their Jacobian values; with M < 3 a mix of -y rows and sample rows.  Fixture = the reference's
    own code executed (tests/golden/make_golden_ref.py)."""
    from riskaversetrajopt_b200.car.driving import Model
    g = np.load(os.path.join(G, "ref_car_relaxed_edge.npz"))
    model = Model(M, method, 0.05, samples=_seed0(M, method))
    for name, it in (("iter0", 0), ("iter1", 1), ("iter0", 0)):
        A, l, u = model.get_constraints_coeffs(g["us1"], it)
        k = f"{method}_M{M}_{name}"
        assert A.shape == tuple(g[k + "_shape"])
        assert np.array_equal(A.indptr, g[k + "_indptr"]) and np.array_equal(A.indices, g[k + "_indices"])
        assert rel_err(A.data, g[k + "_data"]) < RTOL64
        for mine, ref in ((l, g[k + "_l"]), (u, g[k + "_u"])):
            assert np.array_equal(np.isnan(mine), np.isnan(ref)) and np.array_equal(np.isinf(mine), np.isinf(ref))
            f = np.isfinite(ref)
            assert np.allclose(mine[f], ref[f], rtol=RTOL64, atol=1e-11)


def test_baseline_scp_from_iteration_zero():
    """The reference's baseline run: Model(M, 'baseline'), define_problem(us, 0), solve, ... (:536-545)."""
    from riskaversetrajopt_b200.car.driving import Model
    model = Model(10, 'baseline', 0.05, samples=_seed0(10, 'baseline'))
    us = model.initial_guess_us_mat()
    for it in range(3):
        model.define_problem(us, it)
        us, _ = model.solve()
    assert np.all(np.isfinite(us)) and us.shape == (20, 2)


def test_nonfinite_guard_counts_coincident_pedestrians():
    """|p_ego - p_ped| = 0 at k = 0 makes the reference divide by zero (car/driving.py:154) and
    return NaN rows silently; the library counts such samples on the device."""
    from riskaversetrajopt_b200._lib import SaaError
    from riskaversetrajopt_b200.car.driving import Model
    s = [x.copy() for x in _seed0(40)]
    model = Model(40, 'saa', 0.05, samples=tuple(s))
    us = model.initial_guess_us_mat()
    model.get_constraints_coeffs(us, 1)
    assert model.path.check_finite() == 0
    s[0][5, 4:6] = s[0][5, 0:2]                 # pedestrian 5 starts on top of the ego car
    s[0][17, 4:6] = s[0][17, 0:2]
    bad = Model(40, 'saa', 0.05, samples=tuple(s))
    A, l, u = bad.get_constraints_coeffs(us, 1)
    assert np.isnan(A.data).any()
    assert bad.path.check_finite(raise_error=False) == 2
    assert bad.path.check_finite() == 0         # the counter is cleared by the read
    bad.get_constraints_coeffs(us, 1)
    with pytest.raises(SaaError):
        bad.path.check_finite()


@pytest.mark.parametrize("S", [4, 12, 31])
@pytest.mark.parametrize("method", ["saa", "baseline"])
def test_other_horizons_run_the_generic_kernels(S, method):
    """driving_params.S is a module constant in the reference (car/driving_params.py:11); a horizon other
    than 20 runs the generic kernel (csrc/generic_kernels.cuh).  Driven through DevicePath."""
    from oracle.oracle_b import CarOracleB
    from riskaversetrajopt_b200 import _lib
    from riskaversetrajopt_b200.car import driving_params as cp
    from riskaversetrajopt_b200.car.driving import BETA
    from riskaversetrajopt_b200.device_path import DevicePath
    M = 41
    rs = np.random.RandomState(S)
    x0 = np.repeat(np.asarray(cp.state_init, dtype=np.float64)[None, :], M, axis=0)
    x0[:, 4:] += rs.randn(M, 4) * np.array([0.1, 0.1, 1e-2, 1e-2])
    ws, wr = rs.uniform(0.025, 0.175, M), rs.uniform(0.005, 0.095, M)
    DWs = np.sqrt(cp.dt) * rs.randn(M, S, 8)
    if method == 'baseline':
        DWs, ws, wr = 0 * DWs, 0 * ws, 0 * wr
    p = DevicePath(_lib.SAA_CAR, method, S, 0.05, M)
    p.set_params_car(cp, BETA, cp.OSQP_TOL)
    p.set_samples_car(x0, ws, wr, DWs)
    ref = CarOracleB(x0, ws, wr, DWs, method, 0.05)
    us = 0.01 + 0.3 * rs.randn(S, 2)
    for it in (0, 1, 3):
        _check(*p.csc(us, it), *ref.get_constraints_coeffs(us, it))
    assert np.allclose(p.rollout(us).cpu().numpy(), ref.rollout(us), rtol=1e-11, atol=1e-12)
    Z, out3 = p.cvar_terms(us, t_risk=-1.0)
    Zr = ref.monte_carlo_constraints(us)[1]
    assert np.allclose(Z.cpu().numpy(), Zr, rtol=1e-10, atol=1e-12)
    assert np.isclose(out3[0].item(), np.maximum(Zr + 1.0, 0).sum(), rtol=1e-10)
