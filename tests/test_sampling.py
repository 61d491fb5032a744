"""Input contract: the RNG call order of the reference's samplers."""
import numpy as np

from riskaversetrajopt_b200.drone import drone_params as dp
from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters


def _literal_drone_sampler(M):
    """The reference's loop order written out literally (drone/drone_utils.py:61-93)."""
    masses = np.random.uniform(dp.mass_nom - dp.mass_delta, dp.mass_nom + dp.mass_delta, M)
    obs_Qs = np.zeros((M, dp.n_obs, 3, 3))
    for o in range(dp.n_obs):
        for d in range(3):
            delta = np.random.uniform(-dp.obs_radii_deltas, dp.obs_radii_deltas, M)
            for i in range(M):
                obs_Qs[i, o, d, d] = 1. / (dp.obs_radii[o] + delta[i])**2
    DWs = np.zeros((M, dp.S, dp.n_x))
    for i in range(M):
        for t in range(dp.S):
            DWs[i, t, :] = np.sqrt(dp.dt) * np.random.randn(dp.n_x)
    return DWs, masses, obs_Qs


def test_drone_sampler_matches_literal_loop_order():
    np.random.seed(0)
    a = sample_uncertain_parameters('saa', M=37)
    np.random.seed(0)
    b = _literal_drone_sampler(37)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_drone_baseline_sampler():
    np.random.seed(3)
    DWs, masses, obs_Qs = sample_uncertain_parameters('baseline', M=5)
    assert np.all(DWs == 0) and np.all(masses == dp.mass_nom)
    assert np.allclose(obs_Qs[:, 1, 2, 2], 1 / dp.obs_radii[1]**2)
    # the stream advanced by the M mass draws and the M*S*n_x normals (reference :77-92)
    nxt = np.random.uniform()
    np.random.seed(3)
    np.random.uniform(0, 1, 5); np.random.randn(5, dp.S, dp.n_x)
    assert nxt == np.random.uniform()


def test_params_values():
    assert dp.dt == 2.5 and dp.S == 20 and dp.M == 50 and dp.n_obs == 3
    assert np.array_equal(dp.feedback_gain, -np.hstack([0.05 * np.eye(3), 0.25 * np.eye(3)]))
    from riskaversetrajopt_b200.car import driving_params as cp
    assert cp.dt == 0.5 and cp.n_x == 8 and cp.n_u == 2
    assert abs(cp.min_separation_distance - (0.5 + np.hypot(2.695, 1.663))) < 1e-15
