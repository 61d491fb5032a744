"""Gaussian-linearisation comparator (riskaversetrajopt_b200/gaussian.py, host NumPy with closed-form
Jacobians) against the reference's own drone_gaussian.py / driving_gaussian.py executed through
oracle/refexec (jacfwd of b w.r.t. x, the mass and the omegas; fori_loop covariance recursion)."""
import numpy as np
import pytest

from oracle.refexec import load_script, reference_available
from oracle.refexec import minijax as jnp

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference is not present (GPU box)")


def test_drone_gaussian_matches_reference_execution():
    from riskaversetrajopt_b200.gaussian import DroneGaussian
    m = load_script("drone/drone_gaussian.py")
    ref = m.Model(m.S, 'gaussian', 0.1)
    mine = DroneGaussian(m.S, 0.1)
    us = 0.01 + 0.4 * np.random.RandomState(0).randn(m.S, 3)
    xs = np.asarray(ref.us_to_state_trajectory(jnp.array(us)))
    Sig = np.asarray(ref.us_to_covariance_trajectory(jnp.array(us)))
    assert np.allclose(mine.us_to_state_trajectory(us), xs, rtol=1e-13, atol=1e-14)
    assert np.allclose(mine.us_to_covariance_trajectory(us), Sig, rtol=1e-11, atol=1e-18)
    a_state, a_obs = mine.uniform_risk_allocation()
    g_ref = np.asarray(ref.obstacle_avoidance_constraints(jnp.array(xs), jnp.array(Sig), jnp.array(a_state), jnp.array(a_obs)))
    assert np.allclose(mine.obstacle_avoidance_constraints(xs, Sig, a_state, a_obs), g_ref, rtol=1e-11, atol=1e-13)
    assert Sig[-1, 0, 0] > 0 and np.allclose(Sig[-1], Sig[-1].T)


def test_car_gaussian_matches_reference_execution():
    from riskaversetrajopt_b200.gaussian import CarGaussian
    m = load_script("car/driving_gaussian.py")
    ref = m.Model('gaussian', 0.1)
    mine = CarGaussian(0.1)
    us = 0.01 + 0.3 * np.random.RandomState(1).randn(m.S, 2)
    xs = np.asarray(ref.us_to_state_trajectory(jnp.array(us)))
    Sig = np.asarray(ref.us_to_covariance_trajectory(jnp.array(us)))
    assert np.allclose(mine.us_to_state_trajectory(us), xs, rtol=1e-13, atol=1e-14)
    assert np.allclose(mine.us_to_covariance_trajectory(us), Sig, rtol=1e-10, atol=1e-18)
    alphas = (0.1 / m.S) * np.ones(m.S)
    d_ref = np.asarray(ref.separation_distances_at_all_times(jnp.array(xs), jnp.array(Sig), jnp.array(alphas)))
    assert np.allclose(mine.separation_distances_at_all_times(xs, Sig, alphas), d_ref, rtol=1e-11, atol=1e-13)
