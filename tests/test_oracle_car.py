"""Car: Oracle-A (autodiff restatement) vs Oracle-B (analytic) vs golden vectors; sampler order."""
import os

import numpy as np
import pytest

from oracle.oracle_a import CarOracleA, car_sample_parameters
from oracle.oracle_b import CarOracleB
from riskaversetrajopt_b200.car.driving import sample_uncertain_parameters
from conftest import rel_err

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def car_seed0():
    state = np.random.get_state()
    np.random.seed(0)
    out = sample_uncertain_parameters(50, 'saa')
    np.random.set_state(state)
    return out


def test_sampler_matches_reference_loop_order():
    for method in ('saa', 'baseline'):
        np.random.seed(4); a = car_sample_parameters(23, method); na = np.random.rand()
        np.random.seed(4); b = sample_uncertain_parameters(23, method); nb = np.random.rand()
        assert all(np.array_equal(x, y) for x, y in zip(a, b)) and na == nb


@pytest.mark.parametrize("method", ["saa", "baseline"])
def test_a_equals_b(car_seed0, method):
    s = tuple(x[:9] for x in car_seed0)
    a, b = CarOracleA(*s, method, 0.05), CarOracleB(*s, method, 0.05)
    us = a.initial_guess_us_mat() + 0.3 * np.random.RandomState(3).randn(20, 2)
    for it in ((0, 1, 2) if method == 'saa' else (1,)):
        A, l, u = a.get_constraints_coeffs(us, it)
        B, lb, ub = b.get_constraints_coeffs(us, it)
        assert A.shape == B.shape
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
        assert rel_err(B.data, A.data, 1e-300) < 1e-10
        assert np.array_equal(np.isnan(l), np.isnan(lb)) and np.array_equal(np.isinf(l), np.isinf(lb))
        f = np.isfinite(l)
        assert np.allclose(l[f], lb[f], rtol=1e-12, atol=1e-13) and np.allclose(u, ub, rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("src", ["", "ref_"], ids=["oracle_a", "reference_exec"])
def test_golden(car_seed0, src):
    g = np.load(os.path.join(G, src + "car_M50_saa.npz"))
    b = CarOracleB(*car_seed0, 'saa', 0.05)
    for name, us, it in (("iter0", g["us0"], 0), ("iter1", g["us1"], 1), ("iter2", g["us1"], 2)):
        A, l, u = b.get_constraints_coeffs(us, it)
        assert tuple(g[name + "_shape"]) == A.shape
        assert np.array_equal(g[name + "_indptr"], A.indptr) and np.array_equal(g[name + "_indices"], A.indices)
        assert rel_err(A.data, g[name + "_data"], 1e-300) < 1e-10
        assert np.allclose(u, g[name + "_u"], rtol=1e-11, atol=1e-12)
        assert np.array_equal(np.isnan(l), np.isnan(g[name + "_l"]))
    assert g["iter0_indices"].size == 214 and g["iter2_indices"].size == 423 * 50 + 159
    assert np.allclose(b.rollout(g["us1"])[:3], g["Xs_first3"], rtol=1e-13, atol=1e-13)
    assert np.allclose(b.monte_carlo_constraints(g["us1"])[1], g["Z"], rtol=1e-12, atol=1e-12)


def test_final_block_is_sample_independent(car_seed0):
    """ego states carry no noise and no pedestrian coupling."""
    b = CarOracleB(*(x[:6] for x in car_seed0))
    fdu, val, _, _, _, _ = b.per_sample(np.random.RandomState(0).randn(20, 2))
    assert np.all(fdu == fdu[0]) and np.all(val == val[0])


@pytest.mark.parametrize("M", [3, 7, 50])
def test_car_pattern_bit_exact(built_lib, car_seed0, M):
    from riskaversetrajopt_b200.pattern import csc_pattern
    s = tuple(x[:M] for x in car_seed0)
    us = np.random.RandomState(M).randn(20, 2)
    for method, its in (("saa", (0, 1)), ("baseline", (1,))):
        b = CarOracleB(*s, method, 0.05)
        for it in its:
            A, _, _ = b.get_constraints_coeffs(us, it)
            n_rows, n_cols, indptr, indices = csc_pattern('car', method, 20, M, relaxed_pattern=(it < 1))
            assert (n_rows, n_cols) == A.shape and indices.size == A.nnz
            assert np.array_equal(indptr, A.indptr) and np.array_equal(indices, A.indices)
