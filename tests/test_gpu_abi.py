"""Error conventions of the C ABI (include/saa_b200.h): negative status + message, no exceptions
across the boundary, Python shim raises SaaError (a RuntimeError)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_create_rejects_bad_arguments(built_lib):
    lib = built_lib.lib
    h = C.c_void_p()
    assert lib.saa_create(C.byref(h), 7, 0, 0, 10, 10, 0, 20, 0.1, 64, 0) == -1          # unknown problem
    assert b"problem" in lib.saa_last_error(None)
    assert lib.saa_create(C.byref(h), 0, 0, 0, 10, 10, 0, 33, 0.1, 64, 0) == -1          # horizon out of range
    assert b"S <= 32" in lib.saa_last_error(None)
    assert lib.saa_create(C.byref(h), 0, 0, 0, 10, 10, 0, 2, 0.1, 64, 0) == -1
    assert lib.saa_create(C.byref(h), 0, 0, 0, 10, 5, 0, 20, 0.1, 64, 0) == -1           # M_local > M_global
    assert lib.saa_create(C.byref(h), 0, 0, 0, 10, 10, 0, 20, 0.1, 16, 0) == -1          # precision
    assert lib.saa_create(C.byref(h), 0, 0, 0, 10, 10, 0, 20, 0.1, 64, 99) == -1         # device index
    assert h.value is None


def test_call_order_and_null_checks(built_lib):
    import torch
    lib = built_lib.lib
    h = C.c_void_p()
    assert lib.saa_create(C.byref(h), 0, 0, 0, 8, 8, 0, 20, 0.1, 64, 0) == 0
    us = np.zeros((20, 3))
    buf = torch.zeros(20000, dtype=torch.float64, device='cuda')
    # samples / params not set yet -> SAA_ERR_STATE
    rc = lib.saa_linearize_assemble(h, us.ctypes.data, 2, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), None, None, 1, None)
    assert rc == -3 and b"params" in lib.saa_last_error(h)
    assert lib.saa_linearize_assemble(h, None, 2, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), None, None, 1, None) == -1
    assert lib.saa_set_params_car(h, None) == -1
    p = built_lib.CarParams()
    assert lib.saa_set_params_car(h, C.byref(p)) == -1 and b"not a car" in lib.saa_last_error(h)
    assert lib.saa_hopper_friction(h, 20, us.ctypes.data, buf.data_ptr(), buf.data_ptr(), None, None, None) == -1
    assert lib.saa_set_output_geometry(h, 4, 0) == -1                                   # does not contain the samples
    assert lib.saa_destroy(h) == 0 and lib.saa_destroy(None) == 0


def test_python_shim_raises(drone_seed0):
    from riskaversetrajopt_b200._lib import SaaError
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_risk import Model
    DWs, masses, obs_Qs = (x[:5] for x in drone_seed0)
    with pytest.raises(SaaError):
        Model(40, np.zeros((5, 40, 6)), masses, obs_Qs)                                  # horizon beyond the generic kernels' range
    m = Model(dp.S, DWs, masses, obs_Qs)
    with pytest.raises(ValueError):
        m.get_constraints_coeffs(np.zeros((20, 2)), 2)                                   # wrong us shape
    bad = obs_Qs.copy(); bad[0, 0, 0, 1] = 0.3
    with pytest.raises(ValueError):
        Model(dp.S, DWs, masses, bad)                                                    # non-diagonal Q
    assert issubclass(SaaError, RuntimeError)


def test_qp_entry_points_reject_what_they_cannot_solve(built_lib, drone_seed0):
    """saa_qp_*: the CVaR ('saa') program in FP64 only; NULL arguments and bad kinds are status codes."""
    import torch
    from riskaversetrajopt_b200 import _lib
    from riskaversetrajopt_b200.device_path import DevicePath
    from riskaversetrajopt_b200.device_qp import DeviceQP
    lib = built_lib.lib
    buf = np.zeros(1 << 16, dtype=np.int64)
    nb, pl = C.c_int64(), C.c_int64()
    base = DevicePath(_lib.SAA_DRONE, 'baseline', 20, 0.1, 8)
    assert lib.saa_qp_layout(base.handle, buf.ctypes.data, buf.size) == -1 and b"saa" in lib.saa_last_error(base.handle)
    f32 = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, 8, precision='fp32')
    assert lib.saa_qp_partials(f32.handle, 2, C.byref(nb), C.byref(pl)) == -1 and b"FP64" in lib.saa_last_error(f32.handle)
    with pytest.raises(ValueError):
        DeviceQP(f32)
    p = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, 8)
    assert lib.saa_qp_layout(p.handle, buf.ctypes.data, 10) == -1 and b"too small" in lib.saa_last_error(p.handle)
    assert lib.saa_qp_layout(p.handle, None, 10) == -1
    assert lib.saa_qp_partials(p.handle, 7, C.byref(nb), C.byref(pl)) == -1
    assert lib.saa_qp_partials(p.handle, 2, C.byref(nb), C.byref(pl)) == 0 and pl.value == 64 and nb.value >= 1
    assert lib.saa_qp_layout(p.handle, buf.ctypes.data, buf.size) == 0
    assert list(buf[:5]) == [6, 60, 60, 20, 3] and buf[17] == 38 and buf[18] == 1140      # n_fin nu R S blk ... nact nnzJ
    d = torch.zeros(64, dtype=torch.float64, device='cuda')
    assert lib.saa_qp_admm_pass(p.handle, None, None, None, None, 1.0, 0.1, 1e-6, 1.6, None, None, 0, None, None) == -1
    assert lib.saa_qp_reduce(p.handle, d.data_ptr(), 0, 64, 0, d.data_ptr(), None) == -1
    assert lib.saa_qp_dense_step(p.handle, None, d.data_ptr(), 0, None) == -1
    car = DevicePath(_lib.SAA_CAR, 'saa', 20, 0.1, 8)
    assert lib.saa_qp_layout(car.handle, buf.ctypes.data, buf.size) == 0 and list(buf[:5]) == [4, 40, 20, 20, 1]
    assert buf[17] == 38 and buf[18] == 380
    for x in (base, f32, p, car):
        x.close()
