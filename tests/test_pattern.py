"""The closed-form CSC pattern (host code inside libsaa_b200.so) must be bit-exact
with what SciPy derives from the oracle's assembled matrix."""
import numpy as np
import pytest

from oracle.oracle_b import DroneOracleB
from riskaversetrajopt_b200.drone import drone_params as dp


def test_library_exports_every_declared_symbol(built_lib):
    import re, os
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "saa_b200.h")).read()
    declared = set(re.findall(r"\b(saa_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"saa_handle"}
    assert declared == set(built_lib.EXPORTS), declared ^ set(built_lib.EXPORTS)
    for name in declared:
        assert hasattr(built_lib.lib, name)
    assert built_lib.lib.saa_version() >= 100


@pytest.mark.parametrize("method", ["saa", "baseline"])
@pytest.mark.parametrize("M", [1, 2, 7, 50])
def test_drone_pattern_bit_exact(built_lib, drone_seed0, method, M):
    from riskaversetrajopt_b200.pattern import csc_pattern
    DWs, masses, obs_Qs = (x[:M] for x in drone_seed0)
    us = np.random.RandomState(M).randn(dp.S, dp.n_u)
    A, l, u = DroneOracleB(dp.S, DWs, masses, obs_Qs, method, 0.1).get_constraints_coeffs(us, 2)
    n_rows, n_cols, indptr, indices = csc_pattern('drone', method, dp.S, M)
    assert (n_rows, n_cols) == A.shape and indices.size == A.nnz
    assert indptr.dtype == A.indptr.dtype == np.int32
    assert np.array_equal(indptr, A.indptr) and np.array_equal(indices, A.indices)


def test_pattern_sizes_closed_form(built_lib):
    from riskaversetrajopt_b200.pattern import pattern_sizes
    for M in (50, 10**4, 10**6, 10**7):
        assert pattern_sizes('drone', 'saa', 20, M) == (68 + 61 * M, 62 + M, 1263 * M + 180)
        assert pattern_sizes('car', 'saa', 20, M) == (46 + 21 * M, 42 + M, 423 * M + 159)


def test_int64_pattern_when_nnz_exceeds_int32(built_lib):
    from riskaversetrajopt_b200.pattern import csc_pattern, pattern_sizes
    # same small matrix through both index widths
    a = csc_pattern('drone', 'saa', 20, 5, index_dtype=np.int32)
    b = csc_pattern('drone', 'saa', 20, 5, index_dtype=np.int64)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert pattern_sizes('drone', 'saa', 20, 2 * 10**6)[2] > 2**31   # int64 required there


def test_bad_arguments_are_reported(built_lib):
    from riskaversetrajopt_b200._lib import SaaError
    from riskaversetrajopt_b200.pattern import pattern_sizes
    with pytest.raises(SaaError):
        pattern_sizes('drone', 'saa', 1, 5)


def test_public_header_is_plain_c():
    """include/saa_b200.h is the drop-in boundary: it must compile as C99 (no C++ or torch types)."""
    import os
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    hdr = os.path.join(os.path.dirname(__file__), "..", "include", "saa_b200.h")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
