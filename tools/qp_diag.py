import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from riskaversetrajopt_b200.drone import drone_params as dp
from riskaversetrajopt_b200.drone.drone_risk import Model
from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
from riskaversetrajopt_b200.device_qp import DeviceQP
M = int(sys.argv[1])
np.random.seed(0)
DW0, m0, Q0 = sample_uncertain_parameters('saa', M=1000)
reps = M // 1000
DWs = np.sqrt(2.5) * np.random.randn(M, 20, 6); masses = np.random.uniform(29, 35, M); obs_Qs = np.tile(Q0, (reps, 1, 1, 1))
model = Model(dp.S, DWs, masses, obs_Qs, 'saa', 0.1)
P, q = model.get_objective_coeffs()
us = model.initial_guess_us_mat()
dq = DeviceQP(model.path, eps_abs=1e-3, eps_rel=1e-3, max_iter=int(sys.argv[2]))
t0 = time.perf_counter(); dq.setup(P, q, model.path.assemble(us, 2)); torch.cuda.synchronize(); print("setup ms", (time.perf_counter() - t0) * 1e3)
dq.update(model.path.assemble(us, 2))
torch.cuda.synchronize(); t0 = time.perf_counter(); r = dq.solve(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"M={M}: {r.info.iter} iters {dt*1e3:.1f} ms -> {dt/r.info.iter*1e6:.0f} us/iter")
# component timings
import ctypes as C
from riskaversetrajopt_b200._lib import lib, check
def timeit(f, n=50):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): f()
    e1.record(); t_host = (time.perf_counter() - t0) / n; torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, t_host * 1e6
print("pass   us (device, host-issue):", timeit(lambda: dq._launch(2, False)))
print("reduce us:", timeit(lambda: dq._reduced(2, 0)))
red = dq._red[2]
print("dense  us:", timeit(lambda: check(lib.saa_qp_dense_step(model.path.handle, dq.G.data_ptr(), red.data_ptr(), 0, model.path._stream()), model.path.handle)))
print("iterate us:", timeit(lambda: dq._iterate(False)))
print("residuals us:", timeit(lambda: dq._residuals(), 10))
print("gram us:", timeit(lambda: dq._launch(1), 5))
print("scale us:", timeit(lambda: dq._launch(0), 5))
