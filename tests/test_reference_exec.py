"""Reference-pinned parity.

(1) ``ref_*.npz`` were minted by executing the reference's OWN source (tests/golden/
    make_golden_ref.py through oracle/refexec).  Where /root/reference exists (this container;
    not the GPU box) the reference is re-executed here and must reproduce the committed files
    bit for bit in pattern and to rounding in value -- so the fixtures cannot drift from the
    reference.
(2) Every CPU oracle (A: torch autodiff restatement, B: closed forms, C: C/OpenMP port, hopper A/B)
    is checked against those reference-minted fixtures; the GPU tests check the CUDA path against
    the same files (tests/test_gpu_*.py, ``reference_exec`` ids).
"""
import ast
import os

import numpy as np
import pytest

from oracle.refexec import load_script, load_nested, extract_functions, reference_available
from oracle.refexec import minijax as jnp
from conftest import rel_err

G = os.path.join(os.path.dirname(__file__), "golden")
needs_reference = pytest.mark.skipif(not reference_available(),
                                     reason="/root/reference is not present (GPU box)")


def _same_csc(A, l, u, g, key, tol=1e-12):
    assert np.array_equal(A.indptr, g[key + "_indptr"]) and np.array_equal(A.indices, g[key + "_indices"])
    assert rel_err(A.data, g[key + "_data"], 1e-300) < tol
    for mine, ref in ((l, g[key + "_l"]), (u, g[key + "_u"])):
        assert np.array_equal(np.isnan(mine), np.isnan(ref)) and np.array_equal(np.isinf(mine), np.isinf(ref))
        f = np.isfinite(ref)
        assert np.array_equal(np.sign(mine[~f & ~np.isnan(ref)]), np.sign(ref[~f & ~np.isnan(ref)]))
        assert np.allclose(mine[f], ref[f], rtol=tol, atol=1e-14)


# ------------------------------------------------------------------ (1) re-execution
@needs_reference
def test_reexecute_reference_drone_reproduces_fixture():
    m = load_script("drone/drone_risk.py")
    np.random.seed(0)
    DWs, masses, obs_Qs = m.sample_uncertain_parameters('saa', M=m.M)
    model = m.Model(m.S, DWs, masses, obs_Qs, 'saa', 0.1)
    g = np.load(os.path.join(G, "ref_drone_M50_saa_iter2.npz"))
    us = np.asarray(model.initial_guess_us_mat())
    assert np.array_equal(us, g["us"])
    A, l, u = model.get_constraints_coeffs(jnp.array(us), 2)
    assert A.shape == tuple(g["shape"]) == (3118, 112) and A.nnz == 63330
    assert np.array_equal(A.indptr, g["indptr"]) and np.array_equal(A.indices, g["indices"])
    assert np.array_equal(A.data, g["data"]) and np.array_equal(l, g["l"]) and np.array_equal(u, g["u"])


@needs_reference
def test_reexecute_reference_car_edge_cases_reproduce_fixture():
    m = load_script("car/driving.py")
    g = np.load(os.path.join(G, "ref_car_relaxed_edge.npz"))
    for M, method in ((8, 'baseline'), (2, 'saa')):
        m.M = M
        np.random.seed(0)
        model = m.Model(M, method, 0.05)
        for name, it in (("iter0", 0), ("iter1", 1)):
            A, l, u = model.get_constraints_coeffs(jnp.array(g["us1"]), it)
            k = f"{method}_M{M}_{name}"
            assert np.array_equal(A.indptr, g[k + "_indptr"]) and np.array_equal(A.indices, g[k + "_indices"])
            assert np.array_equal(A.data, g[k + "_data"])
            assert np.array_equal(l, g[k + "_l"], equal_nan=True) and np.array_equal(u, g[k + "_u"], equal_nan=True)


@needs_reference
def test_reexecute_reference_hopper_reproduces_fixture():
    m = load_script("hopper/hopper.py")
    g = np.load(os.path.join(G, "ref_hopper_M30.npz"))
    assert np.array_equal(m.intensities, g["intensities"]) and np.array_equal(m.taus, g["taus"])
    for method in ('saa', 'baseline'):
        model = m.Model(m.M, method, 0.2)
        assert np.array_equal(np.asarray(model.slip_risk_constraints(g["Z"])), g[method + "_g"])
        J = np.asarray(m.jacrev(model.slip_risk_constraints)(g["Z"]))
        r, c = np.nonzero(J)
        assert np.array_equal(r, g[method + "_jac_r"]) and np.array_equal(c, g[method + "_jac_c"])
        assert np.array_equal(J[r, c], g[method + "_jac_v"])


@needs_reference
def test_times_variant_and_sampler_come_from_the_reference_files():
    """drone_times.py declares its Model inside ``for M in [20, 30, 50]`` (:70-75): the loader takes
    the class from there; drone_utils.sample_uncertain_parameters is the reference's."""
    mt = load_nested("drone/drone_times.py", ast.For, inject=dict(M=4))
    np.random.seed(3)
    model = mt.Model(4, 'saa', 0.05)
    assert np.asarray(model.DWs).shape == (4, 20, 6) and model.initial_guess_us_mat().shape == (20, 3)
    assert float(np.asarray(model.initial_guess_us_mat())[0, 2]) == 0.01          # :143-144: all three axes
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    m = load_script("drone/drone_risk.py")
    np.random.seed(5); a = m.sample_uncertain_parameters('saa', M=7)
    np.random.seed(5); b = sample_uncertain_parameters('saa', M=7)
    assert all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(a, b))


# ------------------------------------------------------------------ (2) oracles vs reference fixtures
def _car_samples(M, method):
    from riskaversetrajopt_b200.car.driving import sample_uncertain_parameters
    st = np.random.get_state(); np.random.seed(0)
    s = sample_uncertain_parameters(M, method); np.random.set_state(st)
    return s


@pytest.mark.parametrize("M,method", [(50, 'baseline'), (8, 'baseline'), (1, 'saa'), (2, 'saa'), (3, 'saa'),
                                      (1, 'baseline')])
def test_car_oracle_b_relaxation_edge_cases(M, method):
    """car/driving.py:411-415 zeroes rows >= n_x = 8 although only 4 are final rows: with the baseline
    the first four separation rows of sample 0 survive WITH their Jacobian values, with M < 3 a
    mix of -y rows and sample rows does."""
    from oracle.oracle_b import CarOracleB
    g = np.load(os.path.join(G, "ref_car_relaxed_edge.npz"))
    b = CarOracleB(*_car_samples(M, method), method, 0.05)
    for name, it in (("iter0", 0), ("iter1", 1)):
        _same_csc(*b.get_constraints_coeffs(g["us1"], it), g, f"{method}_M{M}_{name}", 1e-10)


@pytest.mark.parametrize("M,method", [(8, 'baseline'), (2, 'saa')])
def test_car_oracle_a_relaxation_edge_cases(M, method):
    from oracle.oracle_a import CarOracleA
    g = np.load(os.path.join(G, "ref_car_relaxed_edge.npz"))
    a = CarOracleA(*_car_samples(M, method), method, 0.05)
    for name, it in (("iter0", 0), ("iter1", 1)):
        _same_csc(*a.get_constraints_coeffs(g["us1"], it), g, f"{method}_M{M}_{name}", 1e-10)


def test_car_sampler_reproduces_the_reference_constructor_draws():
    g = np.load(os.path.join(G, "ref_car_M50_saa.npz"))
    s = _car_samples(50, 'saa')
    assert np.array_equal(s[0], g["states_init"]) and np.array_equal(s[1], g["omegas_speed"])
    assert np.array_equal(s[2], g["omegas_repulsive"])


@pytest.mark.parametrize("method", ["saa", "baseline"])
def test_hopper_oracles_vs_reference_fixture(method):
    from oracle.oracle_hopper import HopperOracleA, HopperOracleB
    from riskaversetrajopt_b200.hopper import hopper as hp
    g = np.load(os.path.join(G, "ref_hopper_M30.npz"))
    st = np.random.get_state(); np.random.seed(1)
    feats = hp.sample_friction_features(hp.M); np.random.set_state(st)
    assert all(np.array_equal(a, g[k]) for a, k in zip(feats, ("intensities", "thetas", "taus")))
    nv = hp.num_vars(hp.M)
    for O in (HopperOracleA, HopperOracleB):
        o = O(hp.M, method, 0.2, *feats)
        assert np.allclose(o.g(g["Z"]), g[method + "_g"], rtol=1e-12, atol=1e-14)
        if hasattr(o, "jac"):
            J = o.jac(g["Z"])
            Jg = np.zeros_like(J); Jg[g[method + "_jac_r"], g[method + "_jac_c"]] = g[method + "_jac_v"]
            assert np.allclose(J, Jg, rtol=1e-12, atol=1e-14)
        if hasattr(o, "hess"):
            H = np.tril(o.hess(g["Z"], g[method + "_lam"]))
            Hg = np.zeros((nv, nv)); Hg[g[method + "_hess_r"], g[method + "_hess_c"]] = g[method + "_hess_v"]
            assert np.allclose(H, Hg, rtol=1e-10, atol=1e-12)


def test_c_port_vs_reference_fixture(drone_seed0):
    """Oracle-C (the OpenMP port used as ``cpu_baseline``) against the reference-minted drone matrix:
    the sample-row entries of the u columns, the sample-row upper bounds and the mean rows."""
    from oracle import cpu_port
    g = np.load(os.path.join(G, "ref_drone_M50_saa_iter2.npz"))
    DWs, masses, obs_Qs = drone_seed0
    M = 50
    Ax, ub, sums, Z, off = cpu_port.drone_assemble(g["us"], masses, DWs, obs_Qs)
    n = Ax.size
    row_s0 = 6 + 1 + M
    rows = g["indices"][:n]
    samp = (rows >= row_s0) & (rows < row_s0 + 60 * M)
    assert samp.sum() == 1140 * M
    assert rel_err(Ax[samp], g["data"][:n][samp], 1e-300) < 1e-10
    assert np.allclose(ub, g["u"][row_s0:row_s0 + 60 * M], rtol=1e-11, atol=1e-13)
    assert np.allclose(sums[-6:] / M, g["l"][:6], rtol=1e-11, atol=1e-13)
    assert np.allclose(Z - 1e-3, g["Z"], rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("M,method", [(50, 'baseline'), (8, 'baseline'), (1, 'saa'), (2, 'saa'), (3, 'saa'),
                                      (1, 'baseline')])
def test_library_relaxed_pattern_equals_the_reference(built_lib, M, method):
    """The closed-form CSC pattern of libsaa_b200 (host code, no GPU) for the car's scp_iter 0
    against what the reference's dense scan produces."""
    from riskaversetrajopt_b200.pattern import csc_pattern
    g = np.load(os.path.join(G, "ref_car_relaxed_edge.npz"))
    for name, relaxed in (("iter0", True), ("iter1", False)):
        k = f"{method}_M{M}_{name}"
        n_rows, n_cols, indptr, indices = csc_pattern('car', method, 20, M, relaxed_pattern=relaxed)
        assert (n_rows, n_cols) == tuple(g[k + "_shape"])
        assert np.array_equal(indptr, g[k + "_indptr"]) and np.array_equal(indices, g[k + "_indices"])
