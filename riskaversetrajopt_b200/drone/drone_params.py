"""Quadrotor problem constants (host side, NumPy).

Drop-in for the reference's ``drone/drone_params.py:1-45``: same module-level
names and values, so scripts written against the reference's params module run
unchanged.  The only difference is the array library: the reference builds the
constants with ``jax.numpy``; this build never imports JAX (the device path is
CUDA behind ``libsaa_b200.so``), so they are plain float64 ``numpy`` arrays.
"""
import numpy as np

# --- solver knobs (reference drone_params.py:3-4) --------------------------
OSQP_POLISH = True
OSQP_TOL = 1e-3

# --- dimensions / horizon (reference drone_params.py:6-12) -----------------
n_x, n_u = 6, 3          # state (p, v) in R^3 x R^3, control = force in R^3
S = 20                   # control switches
M = 50                   # default sample count
T = 50.0                 # horizon [s]
dt = T / S

# --- cost and feedback (reference drone_params.py:13-19) -------------------
R = np.eye(n_u)
feedback_gain = -np.hstack([0.05 * np.eye(n_u), 0.25 * np.eye(n_u)])

# --- physical constants (reference drone_params.py:21-25) ------------------
u_max = 10
mass_nom, mass_delta = 32.0, 3
beta = 1e-2              # diffusion magnitude
drag_coefficient = 0.2

# --- ellipsoidal obstacles (reference drone_params.py:26-41) ---------------
# p is inside obstacle o  <=>  (p - c_o)^T Q_o (p - c_o) <= 1,  Q_o = diag(1/len^2)
obs_positions = np.array([[-1.4, -0.1, 0.0],
                          [-0.7, 0.3, 0.0],
                          [-0.3, 0.25, 0.0]])
obs_radii = np.array([0.3, 0.2, 0.2])
obs_radii_deltas = 0.025
n_obs = obs_positions.shape[0]

# --- boundary conditions (reference drone_params.py:43-44) -----------------
x_init = np.array([-1.9, 0.05, 0.2, 0.0, 0.0, 0.0])
x_final = np.zeros(n_x)
