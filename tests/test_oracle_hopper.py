import numpy as np

from oracle.oracle_hopper import HopperOracleA, HopperOracleB, S, n_x, n_u


def _features(M, seed=1):
    rs = np.random.RandomState(seed)
    I = 0.025 * np.sqrt(2 / 30) * rs.uniform(0, 1, (M, 30))
    return I, rs.uniform(0, np.pi, (M, 30)), rs.uniform(0, 2 * np.pi, (M, 30))


def test_hopper_values_a_equals_b():
    M = 5
    f = _features(M)
    Z = np.random.RandomState(0).randn((S + 1) * n_x + S * n_u + M + 2)
    for method in ('saa', 'baseline'):
        a, b = HopperOracleA(M, method, 0.2, *f), HopperOracleB(M, method, 0.2, *f)
        assert np.allclose(a.g(Z), b.g(Z), rtol=1e-13, atol=1e-14)
    assert a.g(Z).shape == (M * 20,) and HopperOracleA(M, 'saa', 0.2, *f).g(Z).shape == (1 + M + 20 * M + 1,)
