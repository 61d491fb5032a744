"""Static CSC sparsity pattern of the assembled constraint matrix (host only).

Thin wrapper over the handle-free C entry points ``saa_static_pattern_*``; needs
the shared library but no GPU.  Replaces the pattern the reference obtains by
scanning a dense matrix (drone/drone_risk.py:419-420, car/driving.py:417-418).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check

PROBLEMS = {'drone': _lib.SAA_DRONE, 'car': _lib.SAA_CAR}


def pattern_sizes(problem, method, S, M, relaxed_pattern=False):
    r, c, n = C.c_int64(), C.c_int64(), C.c_int64()
    check(lib.saa_static_pattern_sizes(PROBLEMS[problem], _lib.METHODS[method], int(S), int(M),
                                       int(relaxed_pattern), C.byref(r), C.byref(c), C.byref(n)))
    return r.value, c.value, n.value


def csc_pattern(problem, method, S, M, relaxed_pattern=False, index_dtype=None):
    """-> (n_rows, n_cols, indptr, indices).  int32 indices while they fit (what
    SciPy picks for the reference's sizes), int64 beyond."""
    n_rows, n_cols, nnz = pattern_sizes(problem, method, S, M, relaxed_pattern)
    if index_dtype is None:
        index_dtype = np.int32 if max(nnz, n_rows) < 2**31 - 1 else np.int64
    index_dtype = np.dtype(index_dtype)
    indptr = np.empty(n_cols + 1, dtype=index_dtype)
    indices = np.empty(nnz, dtype=index_dtype)
    fn = lib.saa_static_pattern_i32 if index_dtype == np.int32 else lib.saa_static_pattern_i64
    check(fn(PROBLEMS[problem], _lib.METHODS[method], int(S), int(M), int(relaxed_pattern),
             indptr.ctypes.data, indices.ctypes.data))
    return n_rows, n_cols, indptr, indices
