"""Mint ``tests/golden/ref_*.npz`` by EXECUTING THE REFERENCE'S OWN SOURCE.

    python tests/golden/make_golden_ref.py            # needs /root/reference (this container)

The reference cannot be imported as is (it needs jax / osqp / ipyopt / matplotlib, none in the
image).  ``oracle/refexec`` therefore executes the reference's source text -- read from
/root/reference where it lies, nothing copied -- with ``jax`` replaced by ``minijax`` (NumPy array
ops + exact forward-mode dual-number autodiff; Hessians by complex step) and the solver / plotting
imports by inert stubs.  What runs is the reference's ``Model`` code line for line:
``get_constraints_coeffs`` (drone/drone_risk.py:401-423, drone/drone_times.py:410-430,
car/driving.py:399-421) with its ``vmap(jacfwd(...))``, dense ``.at[].set`` packing,
``np.copy`` -> relax -> ``sp.csr_matrix`` -> ``sp.vstack(format='csc')`` tail;
``us_to_state_trajectories``; the Monte-Carlo verification functions; the hopper's
``slip_risk_constraints`` and ``jacrev`` / ``hessian`` of it (hopper/hopper.py:300-367, :568-580).

The files have the same keys as the Oracle-A goldens (make_golden.py), so every parity test runs
against both: the oracle restatement AND the reference's own execution.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.refexec import load_script, load_nested, extract_functions  # noqa: E402
from oracle.refexec import minijax as jnp  # noqa: E402
import ast  # noqa: E402


def _csc(out, key, A, l, u):
    out[key + "_shape"] = np.array(A.shape)
    out[key + "_indptr"], out[key + "_indices"], out[key + "_data"] = A.indptr, A.indices, A.data
    out[key + "_l"], out[key + "_u"] = l, u


def drone(dest=HERE):
    m = load_script("drone/drone_risk.py")
    np.random.seed(0)                                   # drone_risk.py:57
    DWs, masses, obs_Qs = m.sample_uncertain_parameters('saa', M=m.M)    # first repeat, :483-484
    model = m.Model(m.S, DWs, masses, obs_Qs, 'saa', 0.1)
    us = np.asarray(model.initial_guess_us_mat())
    A, l, u = model.get_constraints_coeffs(jnp.array(us), 2)
    Xs = np.asarray(model.us_to_state_trajectories(jnp.array(us)))
    # Monte-Carlo verification function of :656-662 (defined inside ``if B_validate_monte_carlo``)
    ns = dict(m.__dict__, model=model)
    mc = extract_functions("drone/drone_risk.py", ["monte_carlo_no_collisions_constraint_verification"], ns)
    f = mc["monte_carlo_no_collisions_constraint_verification"]
    Z = np.array([float(f(jnp.array(us), masses[i], DWs[i], obs_Qs[i])[1]) for i in range(m.M)])
    np.savez_compressed(os.path.join(dest, "ref_drone_M50_saa_iter2.npz"), us=us, shape=A.shape,
                        indptr=A.indptr, indices=A.indices, data=A.data, l=l, u=u,
                        Xs_first3=Xs[:3], Z=Z)

    # first 8 samples, perturbed iterate, all branches; variant 'times' = the Model of drone_times.py
    rs = np.random.RandomState(123)
    us = us + 0.5 * rs.randn(m.S, m.n_u)
    out = dict(us=us)
    m.M = 8                                              # the methods read the module global M
    mt = load_nested("drone/drone_times.py", ast.For, inject=dict(M=8))
    for method in ('saa', 'baseline'):
        for variant in ('risk', 'times'):
            if variant == 'risk':
                mm = m.Model(m.S, DWs[:8], masses[:8], obs_Qs[:8], method, 0.05)
            else:
                mm = mt.Model(8, method, 0.05)          # draws its own samples (:88-118): overwrite
                mm.masses, mm.obs_Qs, mm.DWs = jnp.array(masses[:8]), jnp.array(obs_Qs[:8]), jnp.array(DWs[:8])
            for it in (0, 2):
                A, l, u = mm.get_constraints_coeffs(jnp.array(us), it)
                k = f"{method}_{variant}_{it}"
                out[k + "_indptr"], out[k + "_indices"] = A.indptr, A.indices
                out[k + "_data"], out[k + "_l"], out[k + "_u"] = A.data, l, u
    np.savez_compressed(os.path.join(dest, "ref_drone_M8_branches.npz"), **out)

    # another horizon: S is a constructor argument of the reference's Model (:70-83)
    S12, M6 = 12, 6
    m.M = M6
    np.random.seed(7)
    DWs, masses, obs_Qs = m.sample_uncertain_parameters('saa', M=M6, S=S12, dt=m.T / S12)
    model = m.Model(S12, DWs, masses, obs_Qs, 'saa', 0.1)
    us = np.asarray(model.initial_guess_us_mat()) + 0.3 * np.random.RandomState(3).randn(S12, 3)
    out = dict(us=us, DWs=DWs, masses=masses, obs_Qs=obs_Qs)
    for it in (0, 2):
        A, l, u = model.get_constraints_coeffs(jnp.array(us), it)
        k = f"iter{it}"
        out[k + "_indptr"], out[k + "_indices"], out[k + "_data"], out[k + "_l"], out[k + "_u"] = \
            A.indptr, A.indices, A.data, l, u
    np.savez_compressed(os.path.join(dest, "ref_drone_S12.npz"), **out)


def _car_model(m, M, method, alpha, seed=0):
    """Model(M, method, alpha) of car/driving.py:83-120 (draws from the global legacy RNG)."""
    m.M = M
    np.random.seed(seed)
    return m.Model(M, method, alpha)


def car(dest=HERE):
    m = load_script("car/driving.py")
    model = _car_model(m, 50, 'saa', 0.05)              # seed 0 (:61), first Model of :472
    us0 = np.asarray(model.initial_guess_us_mat())
    us1 = us0 + 0.4 * np.random.RandomState(7).randn(20, 2)
    out = dict(us0=us0, us1=us1)
    for name, us, it in (("iter0", us0, 0), ("iter1", us1, 1), ("iter2", us1, 2)):
        A, l, u = model.get_constraints_coeffs(jnp.array(us), it)
        _csc(out, name, A, l, u)
    out["Xs_first3"] = np.asarray(model.us_to_state_trajectories(jnp.array(us1)))[:3]
    ns = dict(m.__dict__, model=model)
    f = extract_functions("car/driving.py", ["monte_carlo_separation_constraints_verification"], ns)[
        "monte_carlo_separation_constraints_verification"]
    out["Z"] = np.array([float(f(jnp.array(us1), model.states_init[i], model.omegas_speed[i],
                                 model.omegas_repulsive[i], model.DWs[i])[1]) for i in range(50)])
    out["states_init"] = np.asarray(model.states_init)
    out["omegas_speed"], out["omegas_repulsive"] = np.asarray(model.omegas_speed), np.asarray(model.omegas_repulsive)
    np.savez_compressed(os.path.join(dest, "ref_car_M50_saa.npz"), **out)

    # edge cases of the n_x-based relaxation slice (:411-415): the baseline at scp_iter 0 (what the
    # reference's own driver runs, :536-537) and M < 3
    out = dict(us0=us0, us1=us1)
    for M, method in ((50, 'baseline'), (8, 'baseline'), (1, 'saa'), (2, 'saa'), (3, 'saa'), (1, 'baseline')):
        model = _car_model(m, M, method, 0.05)
        for name, us, it in (("iter0", us1, 0), ("iter1", us1, 1)):
            A, l, u = model.get_constraints_coeffs(jnp.array(us), it)
            _csc(out, f"{method}_M{M}_{name}", A, l, u)
    np.savez_compressed(os.path.join(dest, "ref_car_relaxed_edge.npz"), **out)


def hopper(dest=HERE):
    m = load_script("hopper/hopper.py")                 # draws I, theta, tau with seed 1 (:33, :70-74)
    g0 = np.load(os.path.join(HERE, "hopper_M30.npz"))   # same evaluation point / multipliers as Oracle-A's
    Z = g0["Z"]
    out = dict(Z=Z, intensities=m.intensities, thetas=m.thetas, taus=m.taus)
    for method in ('saa', 'baseline'):
        model = m.Model(m.M, method, 0.2)
        g = np.asarray(model.slip_risk_constraints(Z))
        lam = g0[method + "_lam"]
        J = np.asarray(m.jacrev(model.slip_risk_constraints)(Z))          # :569

        def lagrange_dot_g(x, lagrange):                                   # :575-576
            return jnp.dot(lagrange, model.slip_risk_constraints(x))
        H = np.asarray(m.hessian(lagrange_dot_g)(Z, lam))                  # :578
        r, c = np.nonzero(J)
        hr, hc = np.nonzero(np.tril(H))
        out.update({method + "_g": g, method + "_lam": lam, method + "_jac_r": r, method + "_jac_c": c,
                    method + "_jac_v": J[r, c], method + "_hess_r": hr, method + "_hess_c": hc,
                    method + "_hess_v": H[hr, hc]})
    # Monte-Carlo verification (:901-925) on the model's own 30 samples at Z's trajectory
    model = m.Model(m.M, 'saa', 0.2)
    ns = dict(m.__dict__, model=model)
    f = extract_functions("hopper/hopper.py", ["no_slip_constraint", "no_slip_constraints_verification"], ns)[
        "no_slip_constraints_verification"]
    xs, us = model.convert_z_to_xs_us_mats(Z)
    px = jnp.vmap(model.end_effector_position)(xs)[:, 0]
    px = jnp.concatenate([px[:m.time_jump], px[m.time_land:-1]], axis=0)    # :990-992
    forces = jnp.concatenate([us[:m.time_jump, 2:], us[m.time_land:, 2:]], axis=0)
    out["mc_Z"] = np.array([float(f(px, forces, m.intensities[i], m.thetas[i], m.taus[i])[1])
                            for i in range(m.M)])
    np.savez_compressed(os.path.join(dest, "ref_hopper_M30.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["drone", "car", "hopper"]
    for name in which:
        globals()[name]()
        print("minted ref_" + name)
