"""Uncertain-parameter sampling for the quadrotor problem (input contract).

Mirrors ``sample_uncertain_parameters`` of the reference
(``drone/drone_utils.py:61-93``).  What matters for seed-identical inputs is the
ORDER in which the legacy global ``np.random`` stream is consumed:

1. ``M`` masses                      ~ U(m_nom - dm, m_nom + dm)
2. for each obstacle, for each of 3 axes: ``M`` radius offsets ~ U(-dr, dr)
3. ``M*S*n_x`` standard normals, row-major over (sample, step, state),
   scaled by sqrt(dt)

The reference draws (3) as M*S calls of ``randn(n_x)``; one call of
``randn(M, S, n_x)`` consumes the legacy MT19937 stream identically (checked in
``tests/test_sampling.py`` against a literal double loop).
"""
import numpy as np

from . import drone_params as _p


def sample_uncertain_parameters(method='saa', M=100, S=_p.S, dt=_p.dt):
    """-> (DWs (M,S,n_x), masses (M,), obs_Qs (M,n_obs,3,3)); float64.

    ``method='baseline'`` gives nominal mass/radii and zero noise but still
    consumes ``M`` uniforms for the masses (reference drone_utils.py:77-80)
    and the normals of step 3 (``:87-92``).
    """
    if method not in ('saa', 'baseline'):
        raise ValueError("method must be 'saa' or 'baseline'")
    n_obs, n_x = _p.n_obs, _p.n_x
    spread = _p.mass_delta if method == 'saa' else 0 * _p.mass_delta
    masses = np.random.uniform(_p.mass_nom - spread, _p.mass_nom + spread, M)
    obs_Qs = np.zeros((M, n_obs, 3, 3))
    for o in range(n_obs):
        for d in range(3):
            if method == 'saa':
                length = _p.obs_radii[o] + np.random.uniform(
                    -_p.obs_radii_deltas, _p.obs_radii_deltas, M)
            else:
                length = _p.obs_radii[o]
            obs_Qs[:, o, d, d] = 1. / length**2
    DWs = np.sqrt(dt) * np.random.randn(M, S, n_x)
    if method == 'baseline':
        DWs = 0 * DWs
    return DWs, masses, obs_Qs
