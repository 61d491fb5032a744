"""Oracle-A (autodiff restatement) vs Oracle-B (analytic) vs finite differences."""
import numpy as np
import pytest

from oracle.oracle_a import DroneOracleA, avar_closed_form
from oracle.oracle_b import DroneOracleB
from riskaversetrajopt_b200.drone import drone_params as dp
from conftest import rel_err


@pytest.mark.parametrize("method,variant", [("saa", "risk"), ("baseline", "risk"), ("saa", "times")])
def test_a_equals_b(drone_seed0, method, variant):
    DWs, masses, obs_Qs = (x[:12] for x in drone_seed0)
    a = DroneOracleA(dp.S, DWs, masses, obs_Qs, method, 0.1, variant)
    b = DroneOracleB(dp.S, DWs, masses, obs_Qs, method, 0.1, variant)
    us = a.initial_guess_us_mat() + 0.3 * np.random.RandomState(5).randn(dp.S, dp.n_u)
    for it in (0, 2):
        A, l, u = a.get_constraints_coeffs(us, it)
        B, lb, ub = b.get_constraints_coeffs(us, it)
        assert A.shape == B.shape
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
        assert rel_err(B.data, A.data, 1e-300) < 1e-10
        assert np.array_equal(np.isinf(l), np.isinf(lb))
        fin = np.isfinite(l)
        assert np.allclose(l[fin], lb[fin], rtol=1e-12, atol=1e-13)
        assert np.allclose(u, ub, rtol=1e-11, atol=1e-13)


def test_shapes_and_nnz_match_survey(drone_seed0):
    """(68+61M) x (62+M), nnz = 1263 M + 180, int32 indices."""
    DWs, masses, obs_Qs = drone_seed0
    b = DroneOracleB(dp.S, DWs, masses, obs_Qs, 'saa', 0.1)
    A, l, u = b.get_constraints_coeffs(b_us(), 2)
    M = 50
    assert A.shape == (68 + 61 * M, 62 + M) and A.nnz == 1263 * M + 180
    assert l.shape == u.shape == (68 + 61 * M,)


def b_us():
    us = np.zeros((dp.S, dp.n_u)); us[:, :2] = 0.01
    return us


def test_jacobian_vs_finite_differences(drone_seed0):
    DWs, masses, obs_Qs = (x[:3] for x in drone_seed0)
    b = DroneOracleB(dp.S, DWs, masses, obs_Qs)
    rs = np.random.RandomState(2)
    us = rs.randn(dp.S, dp.n_u)
    _, _, _, g_du, _ = b.per_sample(us)

    def g_of(u):
        Xs, _ = b.rollout(u)
        out = np.empty((3, dp.n_obs, dp.S))
        for o in range(dp.n_obs):
            d = Xs[:, 1:, :2] - dp.obs_positions[o, :2]
            out[:, o] = 1 - obs_Qs[:, o, 0, 0][:, None] * d[..., 0]**2 - obs_Qs[:, o, 1, 1][:, None] * d[..., 1]**2
        return out

    h = 1e-6
    for (t, a) in [(0, 0), (5, 1), (17, 0), (18, 1), (3, 2)]:
        up, um = us.copy(), us.copy()
        up[t, a] += h; um[t, a] -= h
        fd = (g_of(up) - g_of(um)) / (2 * h)
        assert np.allclose(fd, g_du[..., t * 3 + a], rtol=1e-5, atol=1e-6)


def test_monte_carlo_terms(drone_seed0):
    DWs, masses, obs_Qs = (x[:20] for x in drone_seed0)
    a = DroneOracleA(dp.S, DWs, masses, obs_Qs)
    b = DroneOracleB(dp.S, DWs, masses, obs_Qs)
    us = np.random.RandomState(0).randn(dp.S, dp.n_u) * 0.2
    sa, Za = a.monte_carlo_constraints(us)
    sb, Zb = b.monte_carlo_constraints(us)
    assert np.array_equal(sa, sb) and np.allclose(Za, Zb, rtol=1e-12, atol=1e-13)
    assert np.isclose(avar_closed_form(Za, 0.1, 0.2), 0.1 + np.mean(np.maximum(Zb - 0.1, 0)) / 0.2)
