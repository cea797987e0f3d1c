"""Pins the CPU oracle (oracle/fdx_oracle.py) against the reference.

1. every golden vector / known-answer table in the reference's own tests for this path;
2. the committed fixtures produced by running the reference's unmodified source
   (tests/golden/make_golden.py) -- bit for bit in float64 and float32;
3. the reference's analytic-solution checks and error paths.
"""
import numpy as np
import numpy.testing as npt
import pytest

from oracle import fdx_oracle as O
from tests import _golden

# --------------------------------------------------------------- 1. reference's golden vectors
REF_VECTORS = [
    # tests/test_finite_diff.py:40-58
    ("central", 1, [1, 4, 6, 3, 2], [3, 2.5, -0.5, -2, -1]),
    ("central", 2, [1, 4, 6, 3, 2, 4], [3.5, 2.5, -0.5, -2.0, 0.5, 3.5]),
    ("central", 1, [1.2, 1.4], [0.2, 0.2]),
    # :76-108
    ("forward", 1, [1, 4, 6, 3, 2], [3.0, 2.0, -3.0, -1.0, -1.0]),
    ("forward", 2, [1, 4, 6, 3, 2, 4], [3.5, 4.5, -4.0, -2.5, 0.0, 3.5]),
    ("forward", 1, [1.2, 1.4], [0.2, 0.2]),
    # :126-158
    ("backward", 1, [1, 4, 6, 3, 2], [3.0, 3.0, 2.0, -3.0, -1.0]),
    ("backward", 2, [1, 4, 6, 3, 2, 4], [3.5, 4.5, 1.5, -5.5, 0.0, 3.5]),
    ("backward", 1, [1.2, 1.4], [0.2, 0.2]),
    # finite_diff.py:206-208 (docstring)
    ("central", 1, [1.2, 1.3, 2.2, 3.0, 4.5, 5.5, 6.0, 7.0, 8.0, 20.0], [0.1, 0.5, 0.85, 1.15, 1.25, 0.75, 0.75, 1.0, 6.5, 12.0]),
]


@pytest.mark.parametrize("coeff_mode", ["exact", "ref64"])
@pytest.mark.parametrize("method,accuracy,x,expected", REF_VECTORS)
def test_reference_golden_vectors(method, accuracy, x, expected, coeff_mode):
    out = O.difference(np.array(x, dtype=np.float64), accuracy=accuracy, axis=0, derivative=1, method=method, coeff_mode=coeff_mode)
    npt.assert_allclose(out, np.array(expected), rtol=1e-7, atol=1e-12)


def test_reference_coefficient_tables():
    # tests/test_fgrad.py:20-46 (atol=1e-2 there; exact rationals here)
    DF = lambda offs: O.coeffs_exact(offs, 1)
    npt.assert_allclose(DF((0, 1)), [-1.0, 1.0], atol=1e-15)
    npt.assert_allclose(DF((0, 1, 2)), [-3 / 2, 2.0, -1 / 2], atol=1e-15)
    npt.assert_allclose(DF((0, 1, 2, 3)), [-11 / 6, 3.0, -3 / 2, 1 / 3], atol=1e-15)
    npt.assert_allclose(DF((0, 1, 2, 3, 4)), [-25 / 12, 4.0, -3.0, 4 / 3, -1 / 4], atol=1e-15)
    npt.assert_allclose(DF((0, 1, 2, 3, 4, 5)), [-137 / 60, 5.0, -5, 10 / 3, -5 / 4, 1 / 5], atol=1e-15)
    npt.assert_allclose(DF((0, 1, 2, 3, 4, 5, 6)), [-49 / 20, 6.0, -15 / 2, 20 / 3, -15 / 4, 6 / 5, -1 / 6], atol=1e-14)
    npt.assert_allclose(DF((-2.2, 3.2, 5)), [-205 / 972, 280 / 972, -75 / 972], atol=1e-14)
    D0 = lambda offs: O.coeffs_exact(offs, 0)
    npt.assert_allclose(D0((-0.5,)), [1])
    npt.assert_allclose(D0((-0.5, 0.5)), [0.5, 0.5])
    npt.assert_allclose(D0((-0.5, 0.5, 1.5)), [3 / 8, 3 / 4, -1 / 8])
    npt.assert_allclose(D0((-0.5, 0.5, 1.5, 2.5)), [5 / 16, 15 / 16, -5 / 16, 1 / 16])
    # the reference's own arithmetic agrees to its own test tolerance
    npt.assert_allclose(O.coeffs_reference((0, 1, 2, 3, 4, 5, 6), 1, x64=True), DF((0, 1, 2, 3, 4, 5, 6)), atol=1e-2)


def test_survey_coefficients():
    # exact coefficients at the BASELINE configs
    npt.assert_allclose(O.coeffs_exact(O.central_offsets(1, 7), 1), [-1 / 60, 3 / 20, -3 / 4, 0, 3 / 4, -3 / 20, 1 / 60], atol=1e-16)
    npt.assert_allclose(O.coeffs_exact(O.central_offsets(2, 5), 2), [1 / 90, -3 / 20, 3 / 2, -49 / 18, 3 / 2, -3 / 20, 1 / 90], atol=1e-15)
    npt.assert_allclose(O.coeffs_exact(O.forward_offsets(2, 4), 2), [15 / 4, -77 / 6, 107 / 6, -13, 61 / 12, -5 / 6], atol=1e-14)


# --------------------------------------------- 2. fixtures from the reference's own source
@pytest.mark.parametrize("mode,coeff_mode", [("x64", "ref64"), ("x32", "ref32")])
def test_oracle_matches_reference_fixtures_bitwise(mode, coeff_mode):
    cases = _golden.load(mode)
    assert len(cases) > 150
    checked = 0
    for c in cases:
        fn = getattr(O, c["op"])
        if c["error"] is not None:
            with pytest.raises((ValueError, TypeError)):
                fn(c["x"], coeff_mode=coeff_mode, **c["kwargs"])
            continue
        out = fn(c["x"], coeff_mode=coeff_mode, **c["kwargs"])
        assert out.dtype == c["out"].dtype, (c["id"], c["op"])
        assert out.shape == c["out"].shape, (c["id"], c["op"])
        assert np.array_equal(out, c["out"]), (c["id"], c["op"], c["kwargs"], np.abs(out - c["out"]).max())
        checked += 1
    assert checked >= len(cases) - 2


def test_reference_coefficients_vs_exact():
    """golden_coeffs.npz holds what the reference's own solve produces; the oracle's restatement
    of that solve must agree, and the x64 ones must be close to the exact rationals."""
    g = _golden.coeffs()
    worst64 = 0.0
    for key in g.files:
        width, kind, d, a = key.split("_")
        d, a = int(d[1:]), int(a[1:])
        offs = {"central": O.central_offsets(d, a + 1), "forward": O.forward_offsets(d, a), "backward": O.backward_offsets(d, a)}[kind]
        ours = O.coeffs_reference(offs, d, x64=(width == "x64"))
        assert np.array_equal(ours, g[key]), key
        if width == "x64":
            exact = O.coeffs_exact(offs, d)
            worst64 = max(worst64, np.abs(ours - exact).max() / np.abs(exact).max())
    assert worst64 < 1e-6  # measured 3.6e-7: the float64 inverse of the 12-point Vandermonde matrix (d=4, a=8)


# ------------------------------------------------ 3. analytic checks + error paths
def _grid2(n=100, lo=0.0, hi=1.0):
    g = np.linspace(lo, hi, n)
    return g[1] - g[0], np.meshgrid(g, g, indexing="ij")


def test_difference_analytic():
    g = np.linspace(0, 1, 1000)
    dx = g[1] - g[0]
    X, Y = np.meshgrid(g, g, indexing="ij")
    F = np.sin(X) * np.cos(Y)
    npt.assert_allclose(O.difference(F, step_size=dx, axis=0, accuracy=3), np.cos(X) * np.cos(Y), rtol=1e-3)
    npt.assert_allclose(O.difference(F, step_size=dx, axis=1, accuracy=3), -np.sin(X) * np.sin(Y), atol=1e-7)


@pytest.mark.parametrize("method", O.METHODS)
def test_divergence_gradient_laplacian_analytic(method):
    dx, (X, Y) = _grid2()
    F = np.stack([X**2 + Y**3, X**4 + Y**3])
    npt.assert_allclose(O.divergence(F, step_size=(dx, dx), accuracy=7, method=method, keepdims=False), 2 * X + 3 * Y**2, atol=5e-7)
    npt.assert_allclose(O.divergence(F, step_size=(dx, dx), accuracy=7, method=method, keepdims=True), (2 * X + 3 * Y**2)[None], atol=5e-7)
    npt.assert_allclose(O.gradient(X**2 + Y**3, step_size=(dx, dx), accuracy=7, method=method), np.stack([2 * X, 3 * Y**2]), atol=5e-7)
    npt.assert_allclose(O.laplacian(X**4 + Y**3, step_size=(dx, dx), accuracy=10, method=method), 12 * X**2 + 6 * Y, atol=5e-6)


@pytest.mark.parametrize("method", O.METHODS)
def test_curl_analytic(method):
    g = np.linspace(0, 1, 40)
    h = g[1] - g[0]
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    F = np.stack([X**2 + Y**2, X**4 + Y**3, 0 * Z])
    npt.assert_allclose(O.curl(F, step_size=(h, h, h), accuracy=5, method=method), np.stack([0 * X, 0 * Y, 4 * X**3 - 2 * Y]), atol=1e-7)
    g = np.linspace(-1, 1, 50)
    h = g[1] - g[0]
    X, Y = np.meshgrid(g, g, indexing="ij")
    F = np.stack([np.sin(Y), np.cos(X)])
    npt.assert_allclose(O.curl(F, accuracy=4, step_size=h, method=method, keepdims=False), -np.sin(X) - np.cos(Y), atol=1e-4)
    npt.assert_allclose(O.curl(F, accuracy=4, step_size=h, method=method, keepdims=True)[0], -np.sin(X) - np.cos(Y), atol=1e-4)
    with pytest.raises(ValueError):
        O.curl(F[None], accuracy=4, step_size=h, method=method)


@pytest.mark.parametrize("method", O.METHODS)
def test_jacobian_hessian_analytic(method):
    dx, (X, Y) = _grid2(lo=-1.0)
    F = np.stack([X**2 * Y, 5 * X + np.sin(Y)])
    npt.assert_allclose(
        O.jacobian(F, accuracy=4, method=method, step_size=(dx, dx)),
        np.array([[2 * X * Y, X**2], [5 * np.ones_like(X), np.cos(Y)]]),
        atol=1e-7,
    )
    H = O.hessian(X**2 * Y, accuracy=4, step_size=(dx, dx), method=method)
    npt.assert_allclose(H, np.array([[2 * Y, 2 * X], [2 * X, np.zeros_like(X)]]), atol=1e-7)
    H = O.hessian(np.sin(X) + np.cos(Y), accuracy=4, step_size=(dx, dx), method=method)
    npt.assert_allclose(H, np.array([[-np.sin(X), 0 * X], [0 * X, -np.cos(Y)]]), atol=1e-7)


def test_error_paths():
    # tests/test_finite_diff.py:337-360
    for fn in (O.backward_offsets, O.forward_offsets, O.central_offsets):
        with pytest.raises(ValueError):
            fn(0, 1)
        with pytest.raises(ValueError):
            fn(1, 0)
    with pytest.raises(ValueError):
        O.coeffs_exact([1, 2], 2)
    with pytest.raises(ValueError):
        O.difference(np.ones(2), accuracy=2)
    with pytest.raises(ValueError):
        O.difference(np.ones(20), method="sideways")
    with pytest.raises(AssertionError):
        O.gradient(np.ones((8, 8)), accuracy=(1, 2, 3))
    with pytest.raises(ValueError):
        O.gradient(np.ones((8, 8)), accuracy=1.5)


def test_hessian_3d_key_collision_is_restated():
    """in 3-D the reference's dict key 2 is written by d2/dx1^2 and then overwritten by
    d2/dx0dx2 (finite_diff.py:499,510).  The oracle is bug-compatible; hessian_correct is not."""
    x = np.random.default_rng(3).standard_normal((9, 9, 9))
    Hb = O.hessian(x, accuracy=2)
    Hc = O.hessian_correct(x, accuracy=2)
    assert np.array_equal(Hb[1, 1], Hc[0, 2])  # the collision
    assert not np.allclose(Hb[1, 1], Hc[1, 1])
    assert np.array_equal(Hb[0, 0], Hc[0, 0]) and np.array_equal(Hb[2, 2], Hc[2, 2])


def test_dense_matrix_is_the_operator():
    rng = np.random.default_rng(5)
    x = rng.standard_normal(31)
    for method in O.METHODS:
        D = O.dense_matrix(31, accuracy=3, derivative=2, method=method, step_size=0.5)
        npt.assert_allclose(D @ x, O.difference(x, accuracy=3, derivative=2, method=method, step_size=0.5), rtol=1e-12, atol=1e-12)


def test_torch_cpu_arm_equals_the_numpy_oracle():
    """bench.py's CPU arm (oracle/fdx_oracle_torch.py, the reference graph on torch-CPU) against the NumPy oracle."""
    import torch

    from oracle import fdx_oracle_torch as OT

    rng = np.random.default_rng(21)
    for dtype, tol in ((np.float64, 1e-13), (np.float32, 2e-6)):
        x = rng.standard_normal((20, 22, 24)).astype(dtype)
        f = rng.standard_normal((3, 20, 22, 24)).astype(dtype)
        kw = dict(accuracy=4, step_size=(0.1, 0.2, 0.3))
        for name, arr, extra in (("laplacian", x, {}), ("gradient", x, {}), ("divergence", f, {"keepdims": False}), ("curl", f, {})):
            ref = getattr(O, name)(arr, **kw, **extra)
            got = getattr(OT, name)(torch.from_numpy(arr), **kw, **extra).numpy()
            assert got.shape == ref.shape and got.dtype == ref.dtype
            assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref)), name
        d = OT.difference(torch.from_numpy(x), axis=1, accuracy=3, step_size=0.5, derivative=2, method="forward").numpy()
        assert np.max(np.abs(d - O.difference(x, axis=1, accuracy=3, step_size=0.5, derivative=2, method="forward"))) <= tol * np.max(np.abs(d))
