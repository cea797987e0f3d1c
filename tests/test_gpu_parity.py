"""GPU parity tests: the CUDA path (through the Python drop-in -> ctypes -> C ABI) against the
oracle and the committed fixtures.  Tolerances (north star): fp32 <= 1e-5, fp64 <= 1e-12 on the
conditioning scale of BASELINE.md section 5 (tests/_parity.py)."""
import numpy as np
import pytest
import torch

import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi
from finitediffx_b200 import plan as P
from oracle import fdx_oracle as O
from tests import _golden
from tests._parity import TOL, parity_error

pytestmark = pytest.mark.gpu

OPS = ("difference", "gradient", "jacobian", "divergence", "hessian", "laplacian", "curl")


def _gpu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _call(op, x, **kw):
    out = getattr(fdx, op)(_gpu(x), **kw)
    assert out.is_cuda
    return out.cpu().numpy()


# ------------------------------------------------------- fixtures from the reference's own source
@pytest.mark.parametrize("mode", ["x64", "x32"])
def test_golden_fixtures_with_reference_coefficients(mode):
    """Kernel arithmetic parity: both sides use the reference's own coefficients
    (SURVEY section 9 decision 2(i)), so only summation order / FMA contraction differ."""
    x64 = mode == "x64"
    P.set_coefficient_source(lambda offs, d: O.coeffs_reference(offs, d, x64=x64).astype(np.float64))
    try:
        worst = 0.0
        for c in _golden.load(mode):
            if c["error"] is not None:
                with pytest.raises(ValueError):
                    _call(c["op"], c["x"], **c["kwargs"])
                continue
            out = _call(c["op"], c["x"], **c["kwargs"])
            assert out.dtype == c["out"].dtype and out.shape == c["out"].shape, (c["id"], c["op"])
            err = parity_error(c["op"], c["x"], c["kwargs"], out, c["out"])
            worst = max(worst, err)
            assert err <= TOL[out.dtype], (c["id"], c["op"], c["kwargs"], err)
        print(f"golden {mode}: worst parity error {worst:.3e}")
    finally:
        P.set_coefficient_source(None)


def test_golden_fixtures_with_exact_coefficients():
    """Product configuration (exact float64 coefficients).  The reference's float64 inverse is
    itself up to 3.6e-7 off for 12-point stencils, so the tolerance widens by the measured
    coefficient deviation of each case; small stencils must still meet 1e-12."""
    g = _golden.coeffs()
    for c in _golden.load("x64"):
        if c["error"] is not None:
            continue
        kw = c["kwargs"]
        out = _call(c["op"], c["x"], **kw)
        err = parity_error(c["op"], c["x"], kw, out, c["out"])
        acc = kw.get("accuracy", 2 if c["op"] == "hessian" else 1)
        amax = max(acc) if isinstance(acc, tuple) else acc
        dmax = kw.get("derivative", 2 if c["op"] in ("laplacian", "hessian") else 1)
        dev = 0.0
        for kind in ("central", "forward", "backward"):
            for d in range(1, dmax + 1):
                key = f"x64_{kind}_d{d}_a{amax}"
                if key in g.files:
                    offs = {"central": O.central_offsets(d, amax + 1), "forward": O.forward_offsets(d, amax), "backward": O.backward_offsets(d, amax)}[kind]
                    ex = O.coeffs_exact(offs, d)
                    dev = max(dev, np.abs(g[key] - ex).max() / np.abs(ex).max())
        assert err <= 1e-12 + 8 * dev, (c["id"], c["op"], kw, err, dev)
        if dmax + amax <= 6:
            assert err <= 1e-12, (c["id"], c["op"], kw, err)


# --------------------------------------------------------------------- difference: every path
SHAPES_AXES = [
    ((257,), 0),
    ((64, 96), 0),
    ((64, 96), 1),
    ((50, 33), 0),  # inner not a multiple of the vector width
    ((50, 33), 1),  # contiguous axis, unaligned rows
    ((40, 36, 128), 0),
    ((40, 36, 128), 1),
    ((40, 36, 128), 2),
    ((3, 30, 31, 7), 2),
    ((3, 30, 31, 7), -3),
    ((1, 48, 1), 1),
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("method", O.METHODS)
def test_difference_all_kernels(dtype, method):
    rng = np.random.default_rng(1)
    for shape, axis in SHAPES_AXES:
        for d, a in ((1, 1), (1, 2), (2, 4), (1, 6), (3, 3), (2, 8), (4, 8), (4, 2)):
            L, p = d + a - 1, (d + a) // 2
            if shape[axis] < max(2 * L, p + L):
                continue
            x = rng.standard_normal(shape).astype(dtype)
            kw = dict(axis=axis, accuracy=a, derivative=d, method=method, step_size=0.7)
            out = _call("difference", x, **kw)
            ref = O.difference(x, **kw)
            assert out.dtype == ref.dtype
            err = parity_error("difference", x, kw, out, ref)
            assert err <= TOL[np.dtype(dtype)], (shape, axis, d, a, method, err, _cabi.last_kernel())


def test_difference_config2_sample():
    """BASELINE config 2 at reduced size: 2048^2 fp32, a sample of the sweep, both axes."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2048, 2048)).astype(np.float32)
    for d, a, method in ((1, 2, "central"), (2, 4, "forward"), (3, 6, "backward"), (4, 8, "central"), (1, 8, "forward")):
        for axis in (0, 1):
            kw = dict(axis=axis, accuracy=a, derivative=d, method=method, step_size=0.5)
            err = parity_error("difference", x, kw, _call("difference", x, **kw), O.difference(x, **kw))
            assert err <= 1e-5, (d, a, method, axis, err)


# ------------------------------------------------------------------ vector-calculus operators
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("method", O.METHODS)
def test_operators_vs_oracle(dtype, method):
    rng = np.random.default_rng(2)
    x3 = rng.standard_normal((40, 44, 72)).astype(dtype)
    x2 = rng.standard_normal((70, 96)).astype(dtype)
    f3 = rng.standard_normal((3, 36, 40, 64)).astype(dtype)
    f2 = rng.standard_normal((2, 50, 60)).astype(dtype)
    f4 = rng.standard_normal((4, 30, 32, 36)).astype(dtype)
    cases = [
        ("laplacian", x3, dict(accuracy=4, step_size=1 / 39)),
        ("laplacian", x3, dict(accuracy=(2, 4, 6), step_size=(0.1, 0.2, 0.3))),
        ("laplacian", x3, dict(accuracy=8, step_size=1 / 39)),    # radius 5 (BASELINE config 5 stencil)
        ("laplacian", x3, dict(accuracy=10, step_size=1 / 39)),   # radius 6
        ("laplacian", x3, dict(accuracy=2, step_size=0.5)),       # radius 2
        ("laplacian", x3, dict(accuracy=1, step_size=0.5)),       # radius 1
        ("gradient", x3, dict(accuracy=8, step_size=0.1)),        # radius 4
        ("divergence", f3, dict(accuracy=10, step_size=0.1)),     # radius 5
        ("curl", f3, dict(accuracy=8, step_size=0.1)),            # radius 4
        ("curl", f3, dict(accuracy=1, step_size=0.1)),
        ("laplacian", x2, dict(accuracy=6, step_size=0.3)),
        ("gradient", x3, dict(accuracy=4, step_size=(0.1, 0.2, 0.3))),
        ("gradient", x2, dict(accuracy=(2, 5), step_size=0.5)),
        ("divergence", f3, dict(accuracy=6, step_size=(0.1, 0.2, 0.3), keepdims=False)),
        ("divergence", f4, dict(accuracy=2, step_size=0.5, keepdims=True)),
        ("divergence", f2, dict(accuracy=3, step_size=(0.1, 0.3))),
        ("curl", f3, dict(accuracy=4, step_size=(0.1, 0.2, 0.3))),
        ("curl", f2, dict(accuracy=2, step_size=0.2, keepdims=False)),
        ("jacobian", f2, dict(accuracy=4, step_size=(0.1, 0.2))),
        ("jacobian", f4, dict(accuracy=3, step_size=(0.1, 0.2, 0.3))),
        ("hessian", x2, dict(accuracy=4, step_size=(0.1, 0.2))),
        ("hessian", x3[:20, :22, :24].copy(), dict(accuracy=2, step_size=(0.1, 0.2, 0.3))),
    ]
    for op, x, kw in cases:
        kw = dict(kw, method=method)
        out = _call(op, x, **kw)
        ref = getattr(O, op)(x, **kw)
        assert out.shape == ref.shape and out.dtype == ref.dtype, (op, out.shape, ref.shape)
        err = parity_error(op, x, kw, out, ref)
        assert err <= TOL[np.dtype(dtype)], (op, kw, err)


def test_readme_example_config1():
    """BASELINE config 1: README divergence, (3,100,100,100) fp32, accuracy 6, central."""
    g = np.linspace(0, 1, 100)
    h = float(g[1] - g[0])
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    F = np.stack([np.sin(X) * np.cos(Y), np.cos(Y) * Z**2, np.sin(Z) * X]).astype(np.float32)
    kw = dict(accuracy=6, step_size=(h, h, h), keepdims=False, method="central")
    out = _call("divergence", F, **kw)
    ref = O.divergence(F, **kw)
    assert parity_error("divergence", F, kw, out, ref) <= 1e-5
    # and the analytic answer, like the reference's notebook (fp32 field, so loose)
    true = np.cos(X) * np.cos(Y) - np.sin(Y) * Z**2 + np.cos(Z) * X
    assert np.abs(out - true).max() < 5e-2


def test_analytic_solutions_fp64():
    """The reference's analytic checks (tests/test_finite_diff.py:161-307) through the CUDA path."""
    g = np.linspace(0, 1, 100)
    dx = g[1] - g[0]
    X, Y = np.meshgrid(g, g, indexing="ij")
    for method in O.METHODS:
        F = np.stack([X**2 + Y**3, X**4 + Y**3])
        np.testing.assert_allclose(_call("divergence", F, step_size=(dx, dx), accuracy=7, method=method, keepdims=False), 2 * X + 3 * Y**2, atol=5e-7)
        np.testing.assert_allclose(_call("gradient", X**2 + Y**3, step_size=(dx, dx), accuracy=7, method=method), np.stack([2 * X, 3 * Y**2]), atol=5e-7)
        np.testing.assert_allclose(_call("laplacian", X**4 + Y**3, step_size=(dx, dx), accuracy=10, method=method), 12 * X**2 + 6 * Y, atol=5e-6)
    g3 = np.linspace(0, 1, 48)
    h = g3[1] - g3[0]
    X3, Y3, Z3 = np.meshgrid(g3, g3, g3, indexing="ij")
    F = np.stack([X3**2 + Y3**2, X3**4 + Y3**3, 0 * Z3])
    np.testing.assert_allclose(_call("curl", F, step_size=(h, h, h), accuracy=5), np.stack([0 * X3, 0 * Y3, 4 * X3**3 - 2 * Y3]), atol=1e-7)


# ------------------------------------------------------------------------------------ adjoints
@pytest.mark.parametrize("method", O.METHODS)
def test_adjoint_equals_transposed_matrix(method):
    """grad of <D x, g> w.r.t. x must be D^T g with D the oracle's dense matrix, boundary rows included."""
    rng = np.random.default_rng(3)
    for d, a, n in ((1, 1, 12), (1, 4, 33), (2, 4, 40), (3, 3, 29), (2, 8, 64)):
        D = O.dense_matrix(n, accuracy=a, derivative=d, method=method, step_size=0.5)
        for shape, axis in (((n,), 0), ((n, 20), 0), ((12, n), 1), ((6, n, 8), 1)):
            x = _gpu(rng.standard_normal(shape)).requires_grad_(True)
            gnp = rng.standard_normal(shape)
            out = fdx.difference(x, axis=axis, accuracy=a, derivative=d, method=method, step_size=0.5)
            (out * _gpu(gnp)).sum().backward()
            expect = np.moveaxis(np.tensordot(D.T, np.moveaxis(gnp, axis, 0), axes=1), 0, axis)
            scale = np.moveaxis(np.tensordot(np.abs(D.T), np.abs(np.moveaxis(gnp, axis, 0)), axes=1), 0, axis)
            err = np.max(np.abs(x.grad.cpu().numpy() - expect) / np.where(scale == 0, 1, scale))
            assert err <= 1e-12, (d, a, n, shape, axis, err)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_autograd_dot_product_every_operator(dtype):
    """<A x, g> == <x, A^T g> for every operator (the custom_vjp contract)."""
    gen = torch.Generator(device="cuda").manual_seed(4)
    rnd = lambda *s: torch.randn(*s, generator=gen, device="cuda", dtype=dtype)
    cases = [
        ("laplacian", rnd(36, 40, 48), dict(accuracy=4, step_size=0.1)),
        ("laplacian", rnd(50, 60), dict(accuracy=(2, 6), step_size=(0.1, 0.2), method="forward")),
        ("gradient", rnd(36, 40, 48), dict(accuracy=4, step_size=(0.1, 0.2, 0.3))),
        ("divergence", rnd(3, 36, 40, 48), dict(accuracy=6, step_size=0.1, keepdims=False)),
        ("divergence", rnd(2, 40, 48), dict(accuracy=3, step_size=0.1, method="backward")),
        ("curl", rnd(3, 36, 40, 48), dict(accuracy=4, step_size=(0.1, 0.2, 0.3))),
        ("curl", rnd(2, 40, 48), dict(accuracy=2, step_size=0.1)),
        ("jacobian", rnd(4, 30, 32, 36), dict(accuracy=3, step_size=0.2)),
        ("hessian", rnd(40, 48), dict(accuracy=4, step_size=(0.1, 0.2))),
        ("difference", rnd(30, 64), dict(axis=1, accuracy=5, derivative=3, method="backward")),
    ]
    tol = 2e-4 if dtype == torch.float32 else 1e-11
    for op, x, kw in cases:
        x = x.requires_grad_(True)
        out = getattr(fdx, op)(x, **kw)
        g = rnd(*out.shape)
        (gx,) = torch.autograd.grad((out * g).sum(), x)
        lhs = (out.detach().double() * g.double()).sum().item()
        rhs = (gx.double() * x.detach().double()).sum().item()
        norm = (out.detach().double().abs() * g.double().abs()).sum().item()
        assert abs(lhs - rhs) <= tol * norm, (op, kw, lhs, rhs, norm)


def test_fused_kernels_cover_forward_and_adjoint():
    """The 3-D operators and their vjps must run on the fused single-pass kernels (not the composed
    fallback) for the BASELINE stencils, and agree with the composed axis kernels."""
    import os

    x = torch.randn(48, 40, 64, device="cuda", requires_grad=True)
    f = torch.randn(3, 48, 40, 64, device="cuda", requires_grad=True)
    for acc in (2, 4, 6, 8, 10):
        for name, arg, kw in (("laplacian", x, {}), ("gradient", x, {}), ("divergence", f, {"keepdims": False}), ("curl", f, {})):
            out = getattr(fdx, name)(arg, accuracy=acc, step_size=0.1, **kw)
            assert _cabi.last_kernel().startswith("fused_"), (name, acc, _cabi.last_kernel())
            g = torch.randn_like(out)
            (gx,) = torch.autograd.grad(out, arg, g)
            fused_bwd = _cabi.last_kernel()
            with _env(FDX_FUSED_DISABLE=1):
                out2 = getattr(fdx, name)(arg, accuracy=acc, step_size=0.1, **kw)
                (gx2,) = torch.autograd.grad(out2, arg, g)
            tol = 2e-5
            assert ((out - out2).abs().max() / out2.abs().max()).item() < tol, (name, acc)
            assert ((gx - gx2).abs().max() / gx2.abs().max()).item() < tol, (name, acc)
            # the transposed plans (bands of P + L rows) stay on the fused kernels at every accuracy (round 2: the band
            # tables of the parameter block are sized by the radius)
            assert fused_bwd.startswith("fused_"), (name, acc, fused_bwd)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fused_path_covers_one_sided_methods_and_per_axis_accuracy(dtype):
    """method = "forward" / "backward" (finite_diff.py:118-165) and accuracy tuples (:335-336, :562-563) run on the
    fused single-pass kernels as radius-max stencils with zero taps, forward and vjp, and meet the oracle."""
    rng = np.random.default_rng(31)
    x = rng.standard_normal((40, 44, 72)).astype(dtype)
    f = rng.standard_normal((3, 36, 40, 64)).astype(dtype)
    cases = []
    for method in ("forward", "backward"):
        cases += [("laplacian", x, dict(accuracy=2, step_size=0.5, method=method)),       # d + a - 1 = 3
                  ("laplacian", x, dict(accuracy=4, step_size=(0.1, 0.2, 0.3), method=method)),  # 5
                  ("gradient", x, dict(accuracy=4, step_size=0.1, method=method)),      # 4
                  ("divergence", f, dict(accuracy=6, step_size=0.1, method=method, keepdims=False)),  # 6
                  ("curl", f, dict(accuracy=3, step_size=0.1, method=method))]
    cases += [("laplacian", x, dict(accuracy=(2, 4, 6), step_size=(0.1, 0.2, 0.3))),
              ("laplacian", x, dict(accuracy=(8, 2, 4), step_size=0.2)),
              ("gradient", x, dict(accuracy=(6, 1, 3), step_size=0.2)),
              ("divergence", f, dict(accuracy=(1, 8, 2), step_size=0.2)),
              ("curl", f, dict(accuracy=(2, 6, 4), step_size=(0.1, 0.2, 0.3)))]
    for op, arr, kw in cases:
        xt = _gpu(arr).requires_grad_(True)
        out = getattr(fdx, op)(xt, **kw)
        assert _cabi.last_kernel().startswith("fused_"), (op, kw, _cabi.last_kernel())
        ref = getattr(O, op)(arr, **kw)
        err = parity_error(op, arr, kw, out.detach().cpu().numpy(), ref)
        assert err <= TOL[np.dtype(dtype)], (op, kw, err)
        g = torch.randn_like(out)
        (gx,) = torch.autograd.grad(out, xt, g)
        assert _cabi.last_kernel().startswith("fused_"), (op, kw, "vjp", _cabi.last_kernel())
        with _env(FDX_FUSED_DISABLE=1):
            xt2 = _gpu(arr).requires_grad_(True)
            (gx2,) = torch.autograd.grad(getattr(fdx, op)(xt2, **kw), xt2, g)
        tol = 3e-5 if dtype == np.float32 else 1e-11
        assert ((gx - gx2).abs().max() / gx2.abs().max()).item() < tol, (op, kw)


def test_double_backward_is_the_operator_again():
    x = torch.randn(32, 40, dtype=torch.float64, device="cuda", requires_grad=True)
    g = torch.randn(32, 40, dtype=torch.float64, device="cuda", requires_grad=True)
    out = fdx.laplacian(x, accuracy=2, step_size=0.5)
    (gx,) = torch.autograd.grad((out * g).sum(), x, create_graph=True)
    w = torch.randn_like(gx)
    (gg,) = torch.autograd.grad((gx * w).sum(), g)  # = A w
    assert torch.allclose(gg, fdx.laplacian(w, accuracy=2, step_size=0.5), rtol=1e-12, atol=1e-12)


# --------------------------------------------------------------- host marshalling and errors
def test_host_arrays_roundtrip_and_dtype_rule():
    x = np.random.default_rng(5).standard_normal((20, 24))
    out = fdx.laplacian(x, accuracy=2)
    assert isinstance(out, np.ndarray) and out.dtype == np.float64
    np.testing.assert_allclose(out, O.laplacian(x, accuracy=2), rtol=1e-12, atol=1e-12)
    out32 = fdx.difference(x.astype(np.float32), accuracy=2)
    assert out32.dtype == np.float32
    outi = fdx.difference(np.array([1, 4, 6, 3, 2]), accuracy=1, axis=0, derivative=1)
    assert outi.dtype == np.float32
    np.testing.assert_allclose(outi, [3, 2.5, -0.5, -2, -1])
    np.testing.assert_allclose(fdx.difference(np.array([1.2, 1.4]), accuracy=1), [0.2, 0.2], rtol=1e-12)
    t = fdx.gradient(torch.from_numpy(x), accuracy=2)
    assert isinstance(t, torch.Tensor) and not t.is_cuda and t.shape == (2, 20, 24)
    h = fdx.difference(torch.from_numpy(x).cuda().half(), accuracy=2)
    assert h.dtype == torch.float32
    # non-contiguous input
    xt = torch.from_numpy(x).cuda().t()
    np.testing.assert_allclose(fdx.difference(xt, accuracy=2, axis=1).cpu().numpy(), O.difference(x.T, accuracy=2, axis=1), rtol=1e-12, atol=1e-12)


def test_error_paths():
    x = torch.ones(2, device="cuda")
    with pytest.raises(ValueError):
        fdx.difference(x, accuracy=2)
    with pytest.raises(ValueError):
        fdx.difference(torch.ones(20, device="cuda"), method="sideways")
    with pytest.raises(ValueError):
        fdx.curl(torch.ones(1, 2, 20, 20, device="cuda"), accuracy=4)
    with pytest.raises(AssertionError):
        fdx.gradient(torch.ones(8, 8, device="cuda"), accuracy=(1, 2, 3))
    with pytest.raises(ValueError):
        fdx.gradient(torch.ones(8, 8, device="cuda"), accuracy=1.5)
    with pytest.raises(ValueError):
        fdx.laplacian(torch.ones(12, 14, 19, device="cuda"), accuracy=8)


# ------------------------------------------------------- full-size, size-independent properties
def test_full_size_properties_512cubed_fp32():
    """BASELINE config 3 size: properties that need no oracle run -- exactness on polynomials,
    linearity, and agreement between the fused laplacian and divergence(gradient)-free
    compositions of the axis kernel on sub-bricks checked against the oracle."""
    n = 512
    dev = "cuda"
    h = 1.0 / (n - 1)
    g = torch.linspace(0, 1, n, device=dev, dtype=torch.float32)
    X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
    poly = X * X + 2 * Y * Y - 3 * Z * Z + X * Y
    lap = fdx.laplacian(poly, accuracy=4, step_size=h)
    # exact answer 0 (2 + 4 - 6); the fp32 rounding error is bounded on the conditioning scale
    # S = 3 axes * sum|c_k| * max|x| / h^2  (sum|c_k| = 6.04 for the d=2, a=4 stencil, max|x| = 4)
    assert lap.abs().max().item() < 2e-6 * 3 * 6.04 * 4.0 / h**2
    del poly
    a = torch.randn(n, n, n, device=dev)
    b = torch.randn(n, n, n, device=dev)
    la, lb = fdx.laplacian(a, accuracy=4, step_size=h), fdx.laplacian(b, accuracy=4, step_size=h)
    lab = fdx.laplacian(a + 2 * b, accuracy=4, step_size=h)
    ref = la + 2 * lb
    assert ((lab - ref).abs().max() / ref.abs().max()).item() < 5e-6
    # sub-bricks against the oracle: corners (boundary stencils in all three axes) and the centre
    for sl in [(slice(0, 24),) * 3, (slice(n - 24, n),) * 3, (slice(244, 268),) * 3, (slice(0, 24), slice(244, 268), slice(n - 24, n))]:
        # the oracle needs the stencil reach around the brick: take a padded brick, compare its core
        pad = 8
        ext = tuple(slice(max(s.start - pad, 0), min(s.stop + pad, n)) for s in sl)
        brick = a[ext].cpu().numpy()
        # interior bricks see one-sided stencils at artificial edges; compare only points whose
        # stencils coincide: distance >= pad from an artificial edge
        refb = O.laplacian(brick, accuracy=4, step_size=h)
        core = tuple(slice(s.start - e.start, s.stop - e.start) for s, e in zip(sl, ext))
        got = la[sl].cpu().numpy()
        scale = sum(O.conditioning_scale(brick, axis=k, accuracy=4, derivative=2, method="central", step_size=h) for k in range(3))
        err = np.max(np.abs(got - refb[core]) / scale[core])
        assert err <= 1e-5, (sl, err)


# ------------------------------------------------------------- fused-kernel code paths (round 1)
def _env(**kw):
    import contextlib
    import os

    @contextlib.contextmanager
    def cm():
        old = {k: os.environ.get(k) for k in kw}
        os.environ.update({k: str(v) for k, v in kw.items()})
        _cabi.reload_env()  # the library snapshots the FDX_* switches
        try:
            yield
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
            _cabi.reload_env()

    return cm()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_fused_result_does_not_depend_on_the_z_chunking(dtype):
    """Graded / uniform / explicit z chunks, tiny chunks included, give bit-identical outputs: every
    output plane receives the same taps in the same order whatever chunk it falls in."""
    x = torch.randn(150, 40, 72, device="cuda", dtype=dtype)
    f = torch.randn(3, 150, 40, 72, device="cuda", dtype=dtype)
    calls = {
        "laplacian": lambda: fdx.laplacian(x, accuracy=4, step_size=0.1),
        "gradient": lambda: fdx.gradient(x, accuracy=4, step_size=0.1),
        "divergence": lambda: fdx.divergence(f, accuracy=6, step_size=0.1, keepdims=False),
        "curl": lambda: fdx.curl(f, accuracy=2, step_size=0.1),
    }
    for name, call in calls.items():
        ref = call()
        assert _cabi.last_kernel().startswith("fused_"), (name, _cabi.last_kernel())
        for env in (dict(FDX_FUSED_CHUNK=4), dict(FDX_FUSED_CHUNK=7), dict(FDX_FUSED_CHUNK=64), dict(FDX_FUSED_CHUNK=300),
                    dict(FDX_FUSED_CHUNK=48, FDX_FUSED_CHUNK_SHORT=8), dict(FDX_FUSED_CUTS="5,9,40,3,60")):
            with _env(**env):
                out = call()
            assert _cabi.last_kernel().startswith("fused_"), (name, env)
            assert torch.equal(out, ref), (name, env, (out - ref).abs().max().item())


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_x_band_paths_and_awkward_tiles(dtype):
    """x bands: the register fast path, the warp-cooperative path (forced with FDX_XBAND_SLOW, and taken
    on its own by wide adjoint bands) and shapes whose last tile is narrow or partial, against the oracle."""
    rng = np.random.default_rng(7)
    for shape in ((20, 18, 64), (20, 33, 68), (17, 16, 132), (24, 40, 12), (12, 50, 200)):
        x = rng.standard_normal(shape).astype(dtype)
        f = rng.standard_normal((3,) + shape).astype(dtype)
        for acc in (2, 4, 6):
            for slow in (0, 1):
                with _env(FDX_XBAND_SLOW=slow):
                    for op, arr, kw in (("laplacian", x, dict(accuracy=acc, step_size=0.3)), ("divergence", f, dict(accuracy=acc, step_size=0.3, keepdims=False)),
                                        ("gradient", x, dict(accuracy=acc, step_size=0.3)), ("curl", f, dict(accuracy=acc, step_size=0.3))):
                        out = _call(op, arr, **kw)
                        assert _cabi.last_kernel().startswith("fused_"), (op, shape, acc, _cabi.last_kernel())
                        err = parity_error(op, arr, kw, out, getattr(O, op)(arr, **kw))
                        assert err <= TOL[np.dtype(dtype)], (op, shape, acc, slow, err)


def test_tiny_and_thin_grids_take_the_fused_path():
    """Grids of a few tiles (very short z chunks) and slabs thinner than a tile."""
    rng = np.random.default_rng(8)
    # (520, 24, 64): few tiles and a long z axis -- the very short chunks of tiny grids must not overflow the cut table
    for shape in ((8, 8, 8), (9, 300, 16), (300, 9, 16), (16, 9, 300), (100, 100, 100), (520, 24, 64)):
        x = rng.standard_normal(shape).astype(np.float32)
        kw = dict(accuracy=2, step_size=0.5)
        out = _call("laplacian", x, **kw)
        assert _cabi.last_kernel().startswith("fused_"), (shape, _cabi.last_kernel())
        assert parity_error("laplacian", x, kw, out, O.laplacian(x, **kw)) <= 1e-5, shape


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_specialised_marches_match_the_generic_march_and_the_oracle(dtype):
    """Interior tiles, rim tiles (in-register x bands, tile y bands, corners) and the chunks of z-band
    planes take different code copies of the fused kernels.  On grids with at least three tiles each way
    (and on partial last tiles, which stay generic) they must agree BIT FOR BIT with the generic march
    (FDX_FUSED_NOPLAIN=1), with the merged z-band chunks of the old chunk list, with either launch order,
    and meet the parity bound against the oracle."""
    rng = np.random.default_rng(11)
    for shape, accs in (((40, 56, 200), (2, 4, 6)), ((33, 50, 264), (4,)), ((21, 97, 196), (2, 6))):
        x = rng.standard_normal(shape).astype(dtype)
        f = rng.standard_normal((3,) + shape).astype(dtype)
        for acc in accs:
            for op, arr, kw in (("laplacian", x, dict(accuracy=acc, step_size=0.25)), ("divergence", f, dict(accuracy=acc, step_size=(0.1, 0.2, 0.3), keepdims=False)),
                                ("gradient", x, dict(accuracy=acc, step_size=0.25)), ("curl", f, dict(accuracy=acc, step_size=0.25))):
                out = _call(op, arr, **kw)
                assert _cabi.last_kernel().startswith("fused_"), (op, shape, acc, _cabi.last_kernel())
                err = parity_error(op, arr, kw, out, getattr(O, op)(arr, **kw))
                assert err <= TOL[np.dtype(dtype)], (op, shape, acc, err)
                for env in (dict(FDX_FUSED_NOPLAIN=1), dict(FDX_FUSED_ZB_MERGE=1), dict(FDX_FUSED_RIMFIRST=0), dict(FDX_FUSED_RIMFIRST=1), dict(FDX_FUSED_ZB_LAST=0),
                            dict(FDX_XBAND_NODIRECT=1), dict(FDX_FUSED_CHUNK=5), dict(FDX_FUSED_NOPLAIN=1, FDX_FUSED_ZB_MERGE=1, FDX_FUSED_CHUNK=9)):
                    with _env(**env):
                        other = _call(op, arr, **kw)
                    assert np.array_equal(out, other), (op, shape, acc, env, float(np.abs(out - other).max()))


def test_z_extents_around_the_band_planes():
    """n0 from `all planes are band planes` upwards: the z-band chunks, an empty or one-plane middle part."""
    rng = np.random.default_rng(12)
    for acc in (2, 4):
        p = (2 + acc) // 2
        for n0 in range(2 + acc + p - 1, 2 + acc + p + 9):  # from the smallest axis the reference accepts (p + L)
            x = rng.standard_normal((n0, 48, 72))
            kw = dict(accuracy=acc, step_size=0.5)
            out = _call("laplacian", x, **kw)
            err = parity_error("laplacian", x, kw, out, O.laplacian(x, **kw))
            assert err <= 1e-12, (n0, acc, err, _cabi.last_kernel())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_adjoints_through_every_march(dtype):
    """Transposed plans have wide bands (P + L rows, up to 2L + 1 taps): band rows may straddle tiles and
    their taps may leave the tile (n1 = 20 with 16-row tiles: the right band of the FIRST tile reaches row 19).
    vjp of every fused operator: specialised marches == generic march bit for bit, and both agree with the
    composition of axis kernels (FDX_FUSED_DISABLE) within the parity bound; dot-product identity in float64."""
    torch.manual_seed(5)
    tol = 2e-5 if dtype == torch.float32 else 1e-12
    # (84, 51, 68) fp32: two x tiles, the shifted second one starts at column 4 -- inside the 6-column adjoint band
    for shape in ((33, 20, 132), (24, 20, 200), (40, 56, 200), (21, 97, 196), (19, 36, 72), (84, 51, 68), (30, 24, 36)):
        x = torch.randn(shape, device="cuda", dtype=dtype)
        f = torch.randn((3,) + shape, device="cuda", dtype=dtype)
        for acc in (2, 4):
            for op, arr in (("laplacian", x), ("gradient", x), ("divergence", f), ("curl", f)):
                kw = dict(accuracy=acc, step_size=(0.1, 0.2, 0.3))
                if op == "divergence":
                    kw["keepdims"] = False

                def vjp(seed=3):
                    a = arr.clone().requires_grad_(True)
                    out = getattr(fdx, op)(a, **kw)
                    g = torch.randn(out.shape, device="cuda", dtype=dtype, generator=torch.Generator(device="cuda").manual_seed(seed))
                    (ga,) = torch.autograd.grad(out, a, g)
                    return out.detach(), g, ga

                out, g, ga = vjp()
                assert _cabi.last_kernel().startswith("fused_"), (op, shape, acc, _cabi.last_kernel())
                with _env(FDX_FUSED_NOPLAIN=1):
                    _, _, ga_generic = vjp()
                assert torch.equal(ga, ga_generic), (op, shape, acc, (ga - ga_generic).abs().max().item())
                with _env(FDX_FUSED_DISABLE=1):
                    _, _, ga_axis = vjp()
                scale = ga_axis.abs().max().item()
                assert (ga - ga_axis).abs().max().item() <= tol * 50 * scale, (op, shape, acc, (ga - ga_axis).abs().max().item() / scale)
                if dtype == torch.float64:
                    lhs, rhs = float((out * g).sum()), float((arr * ga).sum())
                    assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs)), (op, shape, acc, lhs, rhs)


# ------------------------------------------------------------------ traced step_size (device values)
def test_step_size_as_a_device_value_needs_no_host_read_and_has_a_cotangent():
    """The reference traces ``step_size`` (finite_diff.py:168): a CUDA-tensor step size must not be fetched by the host
    (the call is captured in a CUDA graph here, where any host read raises) and ``grad`` w.r.t. it must be
    -d / h * <g, term> summed over the terms of its axis."""
    torch.manual_seed(12)
    x = torch.randn(24, 28, 32, device="cuda", dtype=torch.float64)
    f = torch.randn(3, 24, 28, 32, device="cuda", dtype=torch.float64)
    hs = (0.1, 0.25, 0.4)
    hd = torch.tensor(hs, device="cuda", dtype=torch.float64)
    cases = [("laplacian", x, {}), ("gradient", x, {}), ("divergence", f, {"keepdims": False}), ("curl", f, {}), ("jacobian", f, {})]
    for op, arr, kw in cases:
        ref = getattr(fdx, op)(arr, accuracy=4, step_size=hs, **kw)
        got = getattr(fdx, op)(arr, accuracy=4, step_size=tuple(hd[k] for k in range(3)), **kw)
        assert torch.allclose(got, ref, rtol=1e-12, atol=1e-12 * ref.abs().max().item()), op
    # one scalar device step for every axis, and `difference`
    got = fdx.laplacian(x, accuracy=2, step_size=hd[0])
    assert torch.allclose(got, fdx.laplacian(x, accuracy=2, step_size=0.1), rtol=1e-12, atol=1e-9)
    got = fdx.difference(x, axis=1, accuracy=3, derivative=2, step_size=hd[1], method="forward")
    assert torch.allclose(got, fdx.difference(x, axis=1, accuracy=3, derivative=2, step_size=0.25, method="forward"), rtol=1e-12, atol=1e-9)
    # cotangent of the step sizes and of the array in one backward pass
    h = hd.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    out = fdx.laplacian(xr, accuracy=4, step_size=(h[0], h[1], h[2]))
    g = torch.randn_like(out)
    gx, gh = torch.autograd.grad(out, (xr, h), g)
    (gx_ref,) = torch.autograd.grad(fdx.laplacian(xr, accuracy=4, step_size=hs), xr, g)
    assert torch.allclose(gx, gx_ref, rtol=1e-11, atol=1e-11 * gx_ref.abs().max().item())
    for k in range(3):
        term = fdx.difference(x, axis=k, accuracy=4, derivative=2, step_size=hs[k])
        want = -2.0 / hs[k] * float((g * term).sum())
        assert abs(float(gh[k]) - want) <= 1e-10 * abs(want), (k, float(gh[k]), want)
    # capture: a host read of the step size (.item(), .cpu()) would raise inside the capture
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fdx.gradient(x, accuracy=2, step_size=hd[2])  # warm-up: plans are created outside the capture
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            captured = fdx.gradient(x, accuracy=2, step_size=hd[2])
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.allclose(captured, fdx.gradient(x, accuracy=2, step_size=0.4), rtol=1e-12, atol=1e-9)


# ------------------------------------------------------------------------------- fused 2-D operators
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("method", O.METHODS)
def test_fused_2d_operators(dtype, method):
    """2-D laplacian / gradient / divergence / curl / jacobian run as ONE fused launch each (fdx_*2d_*): forward against
    the oracle, vjp against the oracle's dense transposed matrices; grids around the vector / chunk sizes; per-axis
    accuracy (zero-padded taps) included.  Most of the reference's own tests are 2-D (tests/test_finite_diff.py:161-307)."""
    rng = np.random.default_rng(41)
    tol = TOL[np.dtype(dtype)]
    for shape in ((70, 96), (33, 64), (200, 24), (64, 260), (97, 132)):
        x = rng.standard_normal(shape).astype(dtype)
        f = rng.standard_normal((2,) + shape).astype(dtype)
        f3 = rng.standard_normal((3,) + shape).astype(dtype)
        cases = [("laplacian", x, dict(accuracy=4, step_size=(0.1, 0.3))), ("laplacian", x, dict(accuracy=(2, 6), step_size=0.5)),
                 ("gradient", x, dict(accuracy=3, step_size=(0.1, 0.3))), ("divergence", f, dict(accuracy=2, step_size=(0.1, 0.3), keepdims=False)),
                 ("divergence", f3, dict(accuracy=5, step_size=0.2)), ("curl", f, dict(accuracy=4, step_size=(0.1, 0.3))),
                 ("jacobian", f3, dict(accuracy=2, step_size=(0.1, 0.3)))]
        for op, arr, kw in cases:
            kw = dict(kw, method=method)
            acc = kw["accuracy"]
            amax = max(acc) if isinstance(acc, tuple) else acc
            d = 2 if op == "laplacian" else 1
            if method != "central" and d + amax - 1 > 6:
                continue  # reach > 6: the composition of axis kernels
            if min(shape) < 2 * (d + amax):
                continue
            xt = _gpu(arr).requires_grad_(True)
            out = getattr(fdx, op)(xt, **kw)
            assert _cabi.last_kernel().startswith("fused_") and _cabi.last_kernel().endswith(("2d", "2d_tma")), (op, shape, kw, _cabi.last_kernel())
            ref = getattr(O, op)(arr, **kw)
            assert out.shape == ref.shape
            err = parity_error(op, arr, kw, out.detach().cpu().numpy(), ref)
            assert err <= tol, (op, shape, kw, err)
            # vjp == oracle: flatten the operator through dense per-axis matrices
            g = rng.standard_normal(ref.shape).astype(dtype)
            (gx,) = torch.autograd.grad(out, xt, _gpu(g))
            assert _cabi.last_kernel().endswith(("2d", "2d_tma")), (op, shape, kw, "vjp", _cabi.last_kernel())
            hs = kw["step_size"] if isinstance(kw["step_size"], tuple) else (kw["step_size"],) * 2
            accs = acc if isinstance(acc, tuple) else (acc,) * 2
            D = [O.dense_matrix(shape[k], accuracy=accs[k], derivative=d, method=method, step_size=hs[k]) for k in range(2)]
            ap = lambda k, a: np.moveaxis(np.tensordot(D[k].T, np.moveaxis(a, k, 0), axes=1), 0, k)  # noqa: E731
            aa = lambda k, a: np.moveaxis(np.tensordot(np.abs(D[k].T), np.abs(np.moveaxis(a, k, 0)), axes=1), 0, k)  # noqa: E731
            g64 = g.astype(np.float64)
            if op == "laplacian":
                want, sc = ap(0, g64) + ap(1, g64), aa(0, g64) + aa(1, g64)
            elif op == "gradient":
                want, sc = ap(0, g64[0]) + ap(1, g64[1]), aa(0, g64[0]) + aa(1, g64[1])
            elif op == "divergence":
                gg = g64[0] if g64.ndim == 3 else g64
                want = np.stack([ap(0, gg), ap(1, gg)] + [np.zeros(shape)] * (arr.shape[0] - 2))
                sc = np.stack([aa(0, gg), aa(1, gg)] + [np.ones(shape)] * (arr.shape[0] - 2))
            elif op == "curl":  # out = D0 f1 - D1 f0
                want, sc = np.stack([-ap(1, g64[0]), ap(0, g64[0])]), np.stack([aa(1, g64[0]), aa(0, g64[0])])
            else:  # jacobian: out[c, k] = D_k f_c
                want = np.stack([ap(0, g64[c, 0]) + ap(1, g64[c, 1]) for c in range(arr.shape[0])])
                sc = np.stack([aa(0, g64[c, 0]) + aa(1, g64[c, 1]) for c in range(arr.shape[0])])
            e2 = float(np.max(np.abs(gx.double().cpu().numpy() - want) / np.where(sc == 0, 1.0, sc)))
            assert e2 <= tol, (op, shape, kw, "vjp", e2)


def test_fused_2d_large_and_fallbacks():
    """(2048, 2048) fp32 through the fused 2-D kernels against the oracle on strided rows / columns; grids the fused
    kernels do not take (last axis not a multiple of the vector) fall back to the axis kernels with the same results."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((2048, 2048), device="cuda", generator=gen)
    kw = dict(accuracy=4, step_size=(0.1, 0.2))
    out = fdx.laplacian(x, **kw)
    assert _cabi.last_kernel() == "fused_laplacian2d_tma"  # 16 MB: the persistent TMA-ring form
    with _env(FDX_FUSED2D_TMA_DISABLE=1):
        plain = fdx.laplacian(x, **kw)
        assert _cabi.last_kernel() == "fused_laplacian2d"
    assert torch.equal(out, plain)
    with _env(FDX_FUSED2D_DISABLE=1):
        ref = fdx.laplacian(x, **kw)
        assert not _cabi.last_kernel().startswith("fused_laplacian2d")
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 2e-6
    cols = x[:, :64].cpu().numpy()  # full columns: the oracle's artificial edge is the right one, 24 columns away from the compared ones
    oc = O.laplacian(cols, **kw)
    s = sum(O.conditioning_scale(cols, axis=k, accuracy=4, derivative=2, method="central", step_size=kw["step_size"][k]) for k in range(2))
    e = float(np.max(np.abs(out[:, :40].cpu().numpy().astype(np.float64) - oc[:, :40]) / s[:, :40]))
    assert e <= 1e-5, e
    xu = torch.randn((100, 97), device="cuda", generator=gen)  # 97 columns: not a multiple of 4
    o2 = fdx.laplacian(xu, accuracy=2)
    assert not _cabi.last_kernel().endswith("2d")
    assert parity_error("laplacian", xu.cpu().numpy(), dict(accuracy=2), o2.cpu().numpy(), O.laplacian(xu.cpu().numpy(), accuracy=2)) <= 1e-5


# ------------------------------------------------------------------ contiguous axis of big arrays: the TMA row kernel
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_row_tma_kernel_equals_axis_row_bit_for_bit_and_the_oracle(dtype):
    """`difference` along the last axis of big arrays runs the persistent TMA-staged row kernel (fdx_rowtma.cu).  Forced
    onto small awkward shapes (partial last strip, partial last row block, 3-D outer, one row) it must equal the L1-gather
    kernel bit for bit (same fma chains) and meet the parity bound against the oracle, for all three methods."""
    rng = np.random.default_rng(21)
    v = 4 if dtype == np.float32 else 2
    for shape in ((37, 129 * v), (8, 250 * v), (3, 5, 160 * v), (1, 128 * v), (19, 128 * v + v)):
        x = rng.standard_normal(shape).astype(dtype)
        for method, d, a in (("central", 1, 2), ("central", 2, 4), ("central", 1, 8), ("central", 4, 8), ("forward", 1, 1), ("forward", 2, 5),
                             ("forward", 4, 8), ("backward", 1, 3), ("backward", 3, 8)):
            kw = dict(axis=-1, accuracy=a, derivative=d, method=method, step_size=0.37)
            with _env(FDX_ROWTMA_MIN_MB=0):
                out = _call("difference", x, **kw)
                assert _cabi.last_kernel() == "axis_rowtma", (shape, method, d, a, _cabi.last_kernel())
            with _env(FDX_ROWTMA_DISABLE=1):
                other = _call("difference", x, **kw)
                assert _cabi.last_kernel().startswith("axis_row"), _cabi.last_kernel()
            assert np.array_equal(out, other), (shape, method, d, a, float(np.abs(out - other).max()))
            err = parity_error("difference", x, kw, out, O.difference(x, **kw))
            assert err <= TOL[np.dtype(dtype)], (shape, method, d, a, err)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_col_tma_kernel_equals_axis_stream_bit_for_bit_and_the_oracle(dtype):
    """`difference` along a strided axis of big arrays runs the persistent TMA-staged kernel of fdx_coltma.cu.  Forced onto
    small awkward shapes (partial last strip, partial last row block, outer > 1, axis length below one row block) it must
    equal axis_stream bit for bit (same fma chains) and meet the parity bound against the oracle, for all three methods."""
    rng = np.random.default_rng(22)
    v = 4 if dtype == np.float32 else 2
    for shape, axis in (((70, 33 * v), 0), ((3, 50, 40 * v), 1), ((2, 3, 45, 32 * v), 2), ((41, 5, 32 * v), 0), ((20, 64 * v + v), 0)):
        x = rng.standard_normal(shape).astype(dtype)
        for method, d, a in (("central", 1, 2), ("central", 2, 4), ("central", 1, 8), ("central", 4, 8), ("forward", 1, 1), ("forward", 2, 5),
                             ("forward", 4, 8), ("backward", 1, 3), ("backward", 3, 8)):
            if 2 * (d + a) > shape[axis]:  # the reference rejects axes shorter than its three row regions
                continue
            kw = dict(axis=axis, accuracy=a, derivative=d, method=method, step_size=0.37)
            with _env(FDX_COLTMA_MIN_MB=0):
                out = _call("difference", x, **kw)
                assert _cabi.last_kernel() == "axis_coltma", (shape, method, d, a, _cabi.last_kernel())
            with _env(FDX_COLTMA_DISABLE=1):
                other = _call("difference", x, **kw)
                assert _cabi.last_kernel() == "axis_stream", _cabi.last_kernel()
            assert np.array_equal(out, other), (shape, axis, method, d, a, float(np.abs(out - other).max()))
            err = parity_error("difference", x, kw, out, O.difference(x, **kw))
            assert err <= TOL[np.dtype(dtype)], (shape, axis, method, d, a, err)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_fused_2d_tma_form_equals_the_plain_form_bit_for_bit(dtype):
    """Big 2-D grids take the persistent TMA-ring form of the fused 2-D kernels.  Forced onto small awkward shapes
    (partial last strip / row block, main regions smaller than one item) it must equal the plain form bit for bit --
    forward and vjp (transposed plans: wide frames), all operators and methods; the plain form is the one checked
    against the oracle in test_fused_2d_operators."""
    gen = torch.Generator(device="cuda").manual_seed(77)
    v = 4 if dtype == torch.float32 else 2
    for shape in ((70, 33 * v), (45, 32 * v), (130, 70 * v), (20, 40 * v)):
        x = torch.randn(shape, device="cuda", dtype=dtype, generator=gen)
        f = torch.randn((2,) + shape, device="cuda", dtype=dtype, generator=gen)
        for method in O.METHODS:
            for op, arr, kw in (("laplacian", x, dict(accuracy=4, step_size=(0.1, 0.3))), ("laplacian", x, dict(accuracy=(2, 5), step_size=0.5)),
                                ("gradient", x, dict(accuracy=3, step_size=(0.1, 0.3))), ("divergence", f, dict(accuracy=2, step_size=(0.1, 0.3), keepdims=False)),
                                ("curl", f, dict(accuracy=4, step_size=(0.1, 0.3)))):
                kw = dict(kw, method=method)
                acc = kw["accuracy"]
                amax = max(acc) if isinstance(acc, tuple) else acc
                d = 2 if op == "laplacian" else 1
                if (method != "central" and d + amax - 1 > 6) or min(shape) < 2 * (d + amax):
                    continue

                def run():
                    a = arr.clone().requires_grad_(True)
                    out = getattr(fdx, op)(a, **kw)
                    k1 = _cabi.last_kernel()
                    g = torch.ones_like(out) * 0.5 + out.detach() * 0.25
                    (ga,) = torch.autograd.grad(out, a, g)
                    return out.detach(), ga, k1

                with _env(FDX_FUSED2D_TMA_MIN_MB=0):
                    o1, g1, k1 = run()
                with _env(FDX_FUSED2D_TMA_DISABLE=1):
                    o2, g2, k2 = run()
                assert k1.endswith("2d_tma") and k2.endswith("2d"), (op, shape, kw, k1, k2)
                assert torch.equal(o1, o2), (op, shape, kw, (o1 - o2).abs().max().item())
                assert torch.equal(g1, g2), (op, shape, kw, "vjp", (g1 - g2).abs().max().item())


def test_hessian_on_a_big_grid_one_pass_per_entry_no_copies():
    """3-D hessian on a grid large enough for the TMA axis kernels: the repeated entries are stored by the producing
    pass (fdx_axis_apply_dup).  Bit-identical to the small-array path (plain kernels + device copies), and the slots
    follow the reference's H[i][j] = F[i + j] (oracle parity of the operator itself: test_operators_vs_oracle)."""
    x = torch.randn((96, 160, 256), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))  # 15.7 MB
    kw = dict(accuracy=(2, 4, 2), step_size=(0.1, 0.2, 0.3))
    with _env(FDX_HESS_DISABLE=1):  # (the 3-D default is the one-pass fused kernel: test_fused_hessian3d_*)
        h = fdx.hessian(x, **kw)
        assert _cabi.last_kernel() in ("axis_rowtma", "axis_coltma"), _cabi.last_kernel()
        with _env(FDX_ROWTMA_DISABLE=1, FDX_COLTMA_DISABLE=1):
            h0 = fdx.hessian(x, **kw)
            assert _cabi.last_kernel() in ("axis_row", "axis_stream"), _cabi.last_kernel()
    assert torch.equal(h, h0)
    for i in range(3):
        for j in range(3):
            for k in range(3):
                for m in range(3):
                    if i + j == k + m:
                        assert torch.equal(h[i, j], h[k, m])
    sub = x[:40, :48, :64].cpu().numpy()
    ref = O.hessian(sub, **kw)
    sc = np.abs(ref).max()
    got = h[:, :, :20, :24, :32].cpu().numpy()
    assert np.max(np.abs(got - ref[:, :, :20, :24, :32])) <= 2e-4 * sc  # (coarse: the brick's far faces are artificial edges)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fused_hessian3d_vs_oracle(dtype):
    """The one-pass 3-D hessian (fdx_hessian3d: main kernel + boundary shell) against the oracle's restatement of
    finite_diff.py:460-529 (bug-compatible F[ax1 + ax2] table): all three methods, the accuracies the kernel covers, per-axis
    accuracy and step size, shapes that are no multiple of the tile and shapes of a single tile."""
    rng = np.random.default_rng(11)
    cases = [
        ((20, 22, 24), dict(accuracy=2, step_size=(0.1, 0.2, 0.3))),
        ((37, 29, 136), dict(accuracy=2, step_size=0.5)),              # two x tiles in fp32, three in fp64
        ((33, 41, 44), dict(accuracy=4, step_size=(0.1, 0.2, 0.3))),    # reach 2 / 3
        ((26, 30, 40), dict(accuracy=3, step_size=0.25)),
        ((30, 27, 36), dict(accuracy=6, step_size=0.5)),                # reach 3 / 4
        ((30, 27, 36), dict(accuracy=(2, 6, 4), step_size=(0.3, 0.2, 0.1))),
        ((28, 30, 32), dict(accuracy=(4, 2, 2), step_size=1.0)),
        ((70, 19, 20), dict(accuracy=2, step_size=0.1)),                # several z chunks (FDX_HESS_ZC below)
    ]
    for method in ("central", "forward", "backward"):
        for shape, kw in cases:
            acc = kw["accuracy"]
            amax = max(acc) if isinstance(acc, tuple) else acc
            if method != "central" and amax > 3:
                continue  # one-sided stencils of reach > 3: the composition of axis passes (checked below)
            kw = dict(kw, method=method)
            x = rng.standard_normal(shape).astype(dtype)
            with _env(FDX_HESS_ZC=16 if shape[0] == 70 else 0):
                out = _call("hessian", x, **kw)
            assert _cabi.last_kernel() == "fused_hessian3d", (shape, kw, _cabi.last_kernel())
            ref = O.hessian(x, **kw)
            assert out.shape == ref.shape and out.dtype == ref.dtype
            err = parity_error("hessian", x, kw, out, ref)
            assert err <= TOL[np.dtype(dtype)], (shape, kw, err)
            for i in range(3):
                for j in range(3):
                    assert np.array_equal(out[i, j], out[min(i + j, 2), i + j - min(i + j, 2)]), (i, j)
    # not covered by the fused kernel: still the GPU, as a composition
    x = rng.standard_normal((30, 30, 30)).astype(dtype)  # last axis not a multiple of the 16-byte vector (fp32)
    kw = dict(accuracy=8, step_size=0.1)
    out = _call("hessian", x, **kw)
    assert _cabi.last_kernel() != "fused_hessian3d"
    assert parity_error("hessian", x, kw, out, O.hessian(x, **kw)) <= TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fused_hessian2d_vs_oracle(dtype):
    """The one-pass 2-D hessian (fdx_hessian2d: a warp per strip marching down axis 0, band columns by shuffles, band rows as
    items of their own) against the oracle, including the reference's own test shape (tests/test_finite_diff.py:286-307:
    100 x 100, accuracy 4)."""
    rng = np.random.default_rng(12)
    cases = [
        ((100, 100), dict(accuracy=4, step_size=(0.02, 0.03))),
        ((40, 48), dict(accuracy=2, step_size=0.5)),
        ((37, 264), dict(accuracy=4, step_size=(0.1, 0.2))),       # three strips in fp32, the last one anchored
        ((90, 20), dict(accuracy=(2, 6), step_size=0.25)),
        ((33, 36), dict(accuracy=6, step_size=0.5)),
        ((64, 132), dict(accuracy=(3, 2), step_size=(0.3, 0.1))),
        ((300, 16), dict(accuracy=2, step_size=0.1)),               # several chunks of rows
    ]
    for method in ("central", "forward", "backward"):
        for shape, kw in cases:
            acc = kw["accuracy"]
            amax = max(acc) if isinstance(acc, tuple) else acc
            if method != "central" and amax > 3:
                continue
            kw = dict(kw, method=method)
            x = rng.standard_normal(shape).astype(dtype)
            with _env(FDX_HESS2D_WPS=1 if shape[0] == 300 else 32):
                out = _call("hessian", x, **kw)
            assert _cabi.last_kernel() == "fused_hessian2d", (shape, kw, _cabi.last_kernel())
            ref = O.hessian(x, **kw)
            assert out.shape == ref.shape and out.dtype == ref.dtype
            err = parity_error("hessian", x, kw, out, ref)
            assert err <= TOL[np.dtype(dtype)], (shape, kw, err)
            assert np.array_equal(out[0, 1], out[1, 0])
    x = torch.randn((1024, 2048), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    kw = dict(accuracy=4, step_size=(0.1, 0.2))
    h = fdx.hessian(x, **kw)
    assert _cabi.last_kernel() == "fused_hessian2d"
    with _env(FDX_HESS_DISABLE=1):
        h0 = fdx.hessian(x, **kw)
    assert _cabi.last_kernel() != "fused_hessian2d"
    sc = h0.abs().amax(dim=(2, 3), keepdim=True)
    assert ((h - h0).abs() / sc).max().item() <= 2e-5


@pytest.mark.parametrize("n_last", [66, 130, 132, 260])
def test_fused_hessians_last_tile_overlapping_the_right_band(n_last):
    """fp64 strips are 64 columns wide: with n_last % 64 == 2 the tile before the anchored last tile overlaps the right band
    columns (3-4 of them at accuracy 4-6).  They belong to the last tile alone; a regression here shows as main-stencil
    values (zero-filled reads past the row) in one or two columns."""
    rng = np.random.default_rng(n_last)
    for acc in (4, 6):
        x3 = rng.standard_normal((20, 21, n_last))
        kw = dict(accuracy=acc, step_size=(0.1, 0.2, 0.3))
        out = _call("hessian", x3, **kw)
        assert _cabi.last_kernel() == "fused_hessian3d"
        assert parity_error("hessian", x3, kw, out, O.hessian(x3, **kw)) <= TOL[np.dtype(np.float64)], (n_last, acc)
        x2 = rng.standard_normal((40, n_last))
        kw = dict(accuracy=acc, step_size=(0.1, 0.2))
        out = _call("hessian", x2, **kw)
        assert _cabi.last_kernel() == "fused_hessian2d"
        assert parity_error("hessian", x2, kw, out, O.hessian(x2, **kw)) <= TOL[np.dtype(np.float64)], (n_last, acc)


def test_fused_hessian3d_against_the_composition_on_a_big_grid():
    """Fused one-pass hessian == the 7-pass composition of axis kernels within the parity bound (the nesting order of the
    mixed entries differs, the boundary rows are the same dense rows), on a grid of several chunks and tiles; inf in the
    input stays confined to the points whose stencils reach it."""
    x = torch.randn((96, 160, 256), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    kw = dict(accuracy=4, step_size=(0.1, 0.2, 0.3))
    h = fdx.hessian(x, **kw)
    assert _cabi.last_kernel() == "fused_hessian3d"
    with _env(FDX_HESS_DISABLE=1):
        h0 = fdx.hessian(x, **kw)
    assert _cabi.last_kernel() != "fused_hessian3d"
    sc = h0.abs().amax(dim=(2, 3, 4), keepdim=True)
    assert ((h - h0).abs() / sc).max().item() <= 2e-5
    xi = x.clone()
    xi[50, 80, 128] = float("inf")
    hi, hi0 = fdx.hessian(xi, **kw), None
    with _env(FDX_HESS_DISABLE=1):
        hi0 = fdx.hessian(xi, **kw)
    assert torch.equal(torch.isfinite(hi), torch.isfinite(hi0))


def test_big_array_kernels_inside_a_cuda_graph_capture():
    """The persistent TMA kernels draw their work-claim counters from a pool: one pair per stream for eager launches, a
    pair of its own (zeroed by a memset node at every replay) for a launch inside a stream capture.  Captured calls
    run the same kernels, give the same bits and can be replayed -- also next to eager launches on the capture stream."""
    x = torch.randn((2304, 1024), device="cuda", generator=torch.Generator(device="cuda").manual_seed(8))  # 9.4 MB
    kw = dict(accuracy=4, step_size=0.3)
    eager = [fdx.difference(x, axis=0, derivative=2, **kw), fdx.difference(x, axis=1, derivative=1, **kw), fdx.laplacian(x, **kw)]
    names = []
    for call in (lambda: fdx.difference(x, axis=0, derivative=2, **kw), lambda: fdx.difference(x, axis=1, derivative=1, **kw), lambda: fdx.laplacian(x, **kw)):
        call()
        names.append(_cabi.last_kernel())
    assert names == ["axis_coltma", "axis_rowtma", "fused_laplacian2d_tma"], names
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            captured = [fdx.difference(x, axis=0, derivative=2, **kw), fdx.difference(x, axis=1, derivative=1, **kw), fdx.laplacian(x, **kw)]
            assert _cabi.last_kernel() == "fused_laplacian2d_tma", _cabi.last_kernel()
    torch.cuda.current_stream().wait_stream(side)
    for rep in range(3):
        for t in captured:
            t.zero_()
        graph.replay()
        with torch.cuda.stream(side):  # eager launches on the capture stream while the graph runs elsewhere
            again = fdx.difference(x, axis=1, derivative=1, **kw)
        torch.cuda.synchronize()
        for a, b in zip(eager, captured):
            assert torch.equal(a, b), rep
        assert torch.equal(again, eager[1])
