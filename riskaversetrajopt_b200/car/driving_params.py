"""Ego car + pedestrian problem constants (host side, NumPy).

Drop-in for the reference's ``car/driving_params.py:1-42`` (same module-level
names/values, ``numpy`` float64 instead of ``jax.numpy``).

State layout (n_x = 8): (px_ego, py_ego, v_ego, phi_ego, px_ped, py_ped,
vx_ped, vy_ped); control (n_u = 2): (acceleration, yaw rate).
"""
import numpy as np

# --- solver knobs (reference driving_params.py:3-4) ------------------------
OSQP_POLISH = True
OSQP_TOL = 3e-4

# --- dimensions / horizon (reference driving_params.py:6-14) ---------------
n_x, n_u = 8, 2
S = 20
M = 50
T = 10.0
dt = T / S

# --- cost (reference driving_params.py:15) ---------------------------------
R = np.diag(np.array([1.0, 1. / 3.0]))

# --- bounds and uncertain social-force gains (reference :17-21) ------------
u_max = 100
omega_speed_nom, omega_speed_del = 0.1, 0.075
omega_repulsive_nom, omega_repulsive_del = 0.05, 0.045

# --- geometry -> minimal centre distance (reference :22-27) ----------------
ego_width, ego_height = 2.695, 1.663      # Smart car length / width
ped_radius = 0.5
min_separation_distance = ped_radius + np.sqrt(ego_width**2 + ego_height**2)

# --- initial / goal states (reference :28-42) ------------------------------
speed_ped_des = 1.3
speed_ego_init = 4
position_ego_init = np.array([-20., 0.])
position_ped_init = np.array([0., -6.])
velocity_ego_init = np.array([speed_ego_init, 0.])
velocity_ped_init = np.array([0., speed_ped_des])
position_ego_goal = np.array([20., 0.1])
velocity_ego_goal = np.array([4.1, 0.])
state_init = np.concatenate(
    (position_ego_init, velocity_ego_init, position_ped_init, velocity_ped_init),
    axis=-1).astype(np.float64)
variance_ped_initial_state = np.diag(np.array([1e-1, 1e-1, 1e-4, 1e-4])**2)
