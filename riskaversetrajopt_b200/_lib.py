"""ctypes binding of ``libsaa_b200.so`` (C ABI: ``include/saa_b200.h``).

There is no CPU fallback: if the shared library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C
riskaversetrajopt_b200/csrc``) importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SAA_B200_LIB lets kernel experiments load an alternative build of the same ABI
LIB_PATH = os.environ.get("SAA_B200_LIB") or os.path.join(_HERE, "libsaa_b200.so")

SAA_DRONE, SAA_CAR, SAA_HOPPER = 0, 1, 2
SAA_METHOD_SAA, SAA_METHOD_BASELINE = 0, 1
SAA_VARIANT_RISK, SAA_VARIANT_TIMES = 0, 1
METHODS = {'saa': SAA_METHOD_SAA, 'baseline': SAA_METHOD_BASELINE}
VARIANTS = {'risk': SAA_VARIANT_RISK, 'times': SAA_VARIANT_TIMES}


class DroneParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("u_max", C.c_double), ("beta", C.c_double),
                ("drag_coefficient", C.c_double), ("gain_p", C.c_double), ("gain_v", C.c_double),
                ("x_init", C.c_double * 6), ("x_final", C.c_double * 6),
                ("n_obs", C.c_int32), ("obs_positions", (C.c_double * 3) * 3),
                ("osqp_tol", C.c_double)]


class HopperPoint(C.Structure):
    _fields_ = [("n_c", C.c_int32), ("px", C.c_double * 32), ("fx", C.c_double * 32),
                ("fz", C.c_double * 32), ("x2", C.c_double * 32), ("x3", C.c_double * 32),
                ("t_risk", C.c_double), ("slack", C.c_double)]


class QpSampleState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("Dy", "Ey", "Es", "xy", "rloc", "zy", "ly", "zs", "ls")]


class CarParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("u_max", C.c_double), ("beta", C.c_double),
                ("speed_ped_des", C.c_double), ("min_separation_distance", C.c_double),
                ("goal", C.c_double * 4), ("osqp_tol", C.c_double)]


# every symbol include/saa_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_PROTOTYPES = {
    "saa_version": (C.c_int, []),
    "saa_last_error": (C.c_char_p, [_H]),
    "saa_create": (C.c_int, [C.POINTER(_H), C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                             C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int]),
    "saa_destroy": (C.c_int, [_H]),
    "saa_set_params_drone": (C.c_int, [_H, C.POINTER(DroneParams)]),
    "saa_set_params_car": (C.c_int, [_H, C.POINTER(CarParams)]),
    "saa_set_samples_drone": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_set_samples_car": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_pattern_sizes": (C.c_int, [_H, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int64)]),
    "saa_pattern_i32": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p]),
    "saa_pattern_i64": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p]),
    "saa_static_pattern_sizes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int,
                                           C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                           C.POINTER(C.c_int64)]),
    "saa_static_pattern_i32": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int,
                                         C.c_void_p, C.c_void_p]),
    "saa_static_pattern_i64": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int,
                                         C.c_void_p, C.c_void_p]),
    "saa_set_output_geometry": (C.c_int, [_H, C.c_int64, C.c_int64]),
    "saa_write_constants": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "saa_linearize_assemble": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "saa_mean_len": (C.c_int64, [_H]),
    "saa_finalize_means": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "saa_merge_shard": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "saa_factored_sizes": (C.c_int, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "saa_linearize_factored": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_expand_factored": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_void_p]),
    "saa_shared_alloc": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    "saa_shared_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
    "saa_shared_close": (C.c_int, [C.c_int, C.c_void_p]),
    "saa_shared_free": (C.c_int, [C.c_int, C.c_void_p]),
    "saa_rollout": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_cvar_terms": (C.c_int, [_H, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "saa_set_samples_hopper": (C.c_int, [_H, C.c_int32, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "saa_hopper_friction": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_check_finite": (C.c_int, [_H, C.POINTER(C.c_int64), C.c_void_p]),
    "saa_measure_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "saa_reserve_sms": (C.c_int, [_H, C.c_int]),
    "saa_peer_inbox_bytes": (C.c_int64, []),
    "saa_peer_allreduce_finalize": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_uint64, C.c_void_p]),
    "saa_hopper_g": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_hopper_jac": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_hopper_g_jac": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_hopper_hess": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_hopper_cvar_terms": (C.c_int, [_H, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_linearize_means": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_select_tail": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "saa_gather_samples": (C.c_int, [_H, _H, C.c_void_p, C.c_void_p]),
    "saa_select_passes": (C.c_int, [_H]),
    "saa_select_begin": (C.c_int, [_H, C.c_int64, C.c_void_p]),
    "saa_select_pass_hist": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "saa_select_pass_pick": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p]),
    "saa_select_counts": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_select_finish": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "saa_set_active": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_int64]),
    "saa_qp_layout": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "saa_qp_partials": (C.c_int, [_H, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "saa_qp_scale_pass": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_qp_gram_pass": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                   C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_qp_admm_pass": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                   C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "saa_qp_check_pass": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "saa_qp_reduce": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "saa_qp_dense_step": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
}
EXPORTS = tuple(_PROTOTYPES)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it first (python -c 'import __graft_entry__ as g; "
            "g.build()').  riskaversetrajopt_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


class SaaError(RuntimeError):
    pass


def check(rc, handle=None):
    """Raise ``SaaError`` (a RuntimeError) for a negative status, like the
    reference's JAX/SciPy calls raise Python exceptions."""
    if rc != 0:
        msg = lib.saa_last_error(handle)
        raise SaaError(f"libsaa_b200 error {rc}: {msg.decode() if msg else '?'}")
