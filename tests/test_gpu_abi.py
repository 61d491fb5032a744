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
