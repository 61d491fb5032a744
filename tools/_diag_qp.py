import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from riskaversetrajopt_b200.drone import drone_params as dp
from riskaversetrajopt_b200.drone.drone_risk import Model
from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
from riskaversetrajopt_b200.device_qp import DeviceQP
for M, eps in ((1000, 1e-3), (20000, 1e-3), (20000, 1e-4), (200000, 1e-3)):
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M) if M <= 20000 else (np.sqrt(2.5) * np.random.randn(M, 20, 6), np.random.uniform(29, 35, M), None)
    if obs_Qs is None:
        d = np.random.uniform(-.025, .025, (M, 3, 3))
        r = np.array(dp.obs_radii) if hasattr(dp, 'obs_radii') else None
        from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters as sup
        DW0, m0, Q0 = sup('saa', M=1000)
        reps = M // 1000
        DWs = np.sqrt(2.5) * np.random.randn(M, 20, 6); masses = np.random.uniform(29, 35, M); obs_Qs = np.tile(Q0, (reps, 1, 1, 1))
    model = Model(dp.S, DWs, masses, obs_Qs, 'saa', 0.1)
    P, q = model.get_objective_coeffs()
    us = model.initial_guess_us_mat()
    dq = DeviceQP(model.path, eps_abs=eps, eps_rel=eps, max_iter=20000)
    t0 = time.perf_counter(); dq.setup(P, q, model.path.assemble(us, 2)); torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
    print(f"M={M} eps={eps} setup {t_setup*1e3:.1f} ms", flush=True)
    for it in range(6):
        t0 = time.perf_counter(); dq.update(model.path.assemble(us, it)); torch.cuda.synchronize(); t_up = time.perf_counter() - t0
        t0 = time.perf_counter(); r = dq.solve(); torch.cuda.synchronize(); t_s = time.perf_counter() - t0
        print(f"  scp {it}: {r.info.status} iters {r.info.iter} rho {dq.rho:.3g} update {t_up*1e3:.1f} ms solve {t_s*1e3:.1f} ms ({t_s/r.info.iter*1e6:.0f} us/iter) t={r.t:.5f} slack={r.slack:.2e}", flush=True)
        us = model.convert_us_vec_to_us_mat(r.u)
    del model, dq
