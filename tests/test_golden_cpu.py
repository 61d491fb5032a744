"""Oracle-B against the committed golden vectors: those minted from Oracle-A (make_golden.py) and
those minted by executing the reference's own source (ref_*, make_golden_ref.py)."""
import os

import numpy as np
import pytest

from oracle.oracle_b import DroneOracleB
from riskaversetrajopt_b200.drone import drone_params as dp
from conftest import rel_err

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("src", ["", "ref_"], ids=["oracle_a", "reference_exec"])
def test_drone_m50_golden(drone_seed0, src):
    g = np.load(os.path.join(G, src + "drone_M50_saa_iter2.npz"))
    DWs, masses, obs_Qs = drone_seed0
    b = DroneOracleB(dp.S, DWs, masses, obs_Qs, 'saa', 0.1)
    A, l, u = b.get_constraints_coeffs(g["us"], 2)
    assert tuple(g["shape"]) == A.shape
    assert np.array_equal(g["indptr"], A.indptr) and np.array_equal(g["indices"], A.indices)
    assert rel_err(A.data, g["data"], 1e-300) < 1e-10
    assert np.allclose(u, g["u"], rtol=1e-11, atol=1e-13)
    Xs, _ = b.rollout(g["us"])
    assert np.allclose(Xs[:3], g["Xs_first3"], rtol=1e-13, atol=1e-14)
    assert np.allclose(b.monte_carlo_constraints(g["us"])[1], g["Z"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("src", ["", "ref_"], ids=["oracle_a", "reference_exec"])
def test_drone_m8_branches_golden(drone_seed0, src):
    g = np.load(os.path.join(G, src + "drone_M8_branches.npz"))
    DWs, masses, obs_Qs = (x[:8] for x in drone_seed0)
    for method in ('saa', 'baseline'):
        for variant in ('risk', 'times'):
            b = DroneOracleB(dp.S, DWs, masses, obs_Qs, method, 0.05, variant)
            for it in (0, 2):
                A, l, u = b.get_constraints_coeffs(g["us"], it)
                k = f"{method}_{variant}_{it}"
                assert np.array_equal(g[k + "_indptr"], A.indptr)
                assert np.array_equal(g[k + "_indices"], A.indices)
                assert rel_err(A.data, g[k + "_data"], 1e-300) < 1e-10
                assert np.allclose(u, g[k + "_u"], rtol=1e-11, atol=1e-13)
                assert np.array_equal(l, g[k + "_l"]) or np.allclose(l, g[k + "_l"], rtol=1e-11, atol=1e-13)
