"""CPU-only tests: host maths, plans, argument handling, and the C-ABI library's exports."""
import ctypes
import os
import re

import numpy as np
import numpy.testing as npt
import pytest

from finitediffx_b200 import plan as P
from finitediffx_b200 import utils as U
from oracle import fdx_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_offsets_match_oracle_and_reject_bad_arguments():
    for d in range(1, 5):
        for a in range(1, 9):
            assert U._generate_forward_offsets(d, a) == O.forward_offsets(d, a)
            assert U._generate_backward_offsets(d, a) == O.backward_offsets(d, a)
            if a >= 2:
                assert U._generate_central_offsets(d, a) == O.central_offsets(d, a)
    for fn in (U._generate_backward_offsets, U._generate_forward_offsets, U._generate_central_offsets):
        with pytest.raises(ValueError):
            fn(derivative=0, accuracy=1)
        with pytest.raises(ValueError):
            fn(derivative=1, accuracy=0)
    with pytest.raises(ValueError):
        U.generate_finitediff_coeffs(offsets=[1, 2], derivative=2)


def test_coefficients_known_answers():
    DF = lambda offs: U.generate_finitediff_coeffs(offs, 1)
    npt.assert_allclose(DF((0, 1, 2, 3, 4, 5, 6)), [-49 / 20, 6.0, -15 / 2, 20 / 3, -15 / 4, 6 / 5, -1 / 6], rtol=0, atol=1e-15)
    npt.assert_allclose(DF((-2.2, 3.2, 5)), [-205 / 972, 280 / 972, -75 / 972], atol=1e-14)
    npt.assert_allclose(U.generate_finitediff_coeffs((-0.5, 0.5, 1.5, 2.5), 0), [5 / 16, 15 / 16, -5 / 16, 1 / 16])
    npt.assert_array_equal(U.generate_finitediff_coeffs((-1, 0, 1), 1), [-0.5, 0.0, 0.5])
    npt.assert_array_equal(U.generate_finitediff_coeffs((-1, 0, 1), 2), [1.0, -2.0, 1.0])
    # config 5 stencil
    c = U.generate_finitediff_coeffs(U._generate_central_offsets(2, 9), 2)
    npt.assert_allclose(c, [1 / 3150, -5 / 1008, 5 / 126, -5 / 21, 5 / 3, -5269 / 1800, 5 / 3, -5 / 21, 5 / 126, -5 / 1008, 1 / 3150], rtol=1e-15)
    # exact where the reference's own int32/float32 solve is not
    offs = U._generate_forward_offsets(3, 8)
    assert np.abs(O.coeffs_reference(offs, 3, x64=False) - U.generate_finitediff_coeffs(offs, 3)).max() > 1.0


def test_check_and_return():
    assert U._check_and_return(3, 2, "accuracy") == (3, 3)
    assert U._check_and_return((1, 2), 2, "accuracy") == (1, 2)
    assert U._check_and_return(0.5, 3, "step_size") == (0.5, 0.5, 0.5)
    assert U._check_and_return(np.float64(0.25), 2, "step_size") == (0.25, 0.25)
    with pytest.raises(AssertionError):
        U._check_and_return((1, 2, 3), 2, "accuracy")
    with pytest.raises(ValueError):
        U._check_and_return(1.5, 2, "accuracy")
    with pytest.raises(ValueError):
        U._check_and_return("central", 2, "accuracy")


@pytest.mark.parametrize("method", O.METHODS)
def test_plans_reproduce_the_oracle_matrix_and_its_transpose(method):
    for d in (1, 2, 3, 4):
        for a in (1, 2, 3, 4, 6, 8):
            p, L = (d + a) // 2, d + a - 1
            nmin = max(p + L, 2 * p + 1) if method == "central" else 2 * L
            for n in (nmin, nmin + 1, 2 * nmin + 3, 97):
                D = O.dense_matrix(n, accuracy=a, derivative=d, method=method)
                fw = P.axis_plan(n, d, a, method)
                assert np.array_equal(fw.dense(), D)
                assert np.array_equal(P.axis_plan_t(n, d, a, method).dense(), D.T)
                assert fw.taps <= P.MAX_TAPS and P.axis_plan_t(n, d, a, method).taps <= P.MAX_TAPS


def test_plan_edge_and_error_zones():
    npt.assert_array_equal(P.axis_plan(2, 1, 1, "central").dense(), [[-1, 1], [-1, 1]])
    for bad in [(5, 1, 4, "central"), (2, 1, 2, "central"), (7, 1, 4, "forward"), (7, 1, 4, "backward"), (20, 1, 2, "sideways"), (4, 1, 3, "central")]:
        with pytest.raises(ValueError):
            P.axis_plan(*bad)
        # the oracle (and the reference) refuse the same sizes
        with pytest.raises((ValueError, TypeError)):
            O.difference(np.ones(bad[0]), derivative=bad[1], accuracy=bad[2], method=bad[3])


def test_coefficient_source_hook():
    try:
        P.set_coefficient_source(lambda offs, d: O.coeffs_reference(offs, d, x64=True))
        D = O.dense_matrix(40, accuracy=6, derivative=2, method="central", coeff_mode="ref64")
        assert np.array_equal(P.axis_plan(40, 2, 6, "central").dense(), D)
    finally:
        P.set_coefficient_source(None)
    assert np.array_equal(P.axis_plan(40, 2, 6, "central").dense(), O.dense_matrix(40, accuracy=6, derivative=2))


def test_library_exports_every_declared_symbol():
    """Build (nvcc cross-compiles without a GPU), load, and check include/fdx_b200.h against the .so."""
    from finitediffx_b200 import _build, _cabi

    path = _build.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "fdx_b200.h")).read()
    declared = set(re.findall(r"\b(fdx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert _cabi.load().fdx_version() >= 1
    assert _cabi.load().fdx_error_string(-2).decode().startswith("fdx: unsupported")


def test_no_cpu_fallback_without_gpu():
    import torch

    import finitediffx_b200 as fdx
    from finitediffx_b200 import _cabi

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_cabi.FdxError):
        fdx.laplacian(np.ones((8, 8)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "finitediffx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f
