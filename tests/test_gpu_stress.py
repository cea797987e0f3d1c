"""Randomised shapes: the fused 3-D kernels (every march, forward and vjp) against the composition of axis
kernels inside the same library (FDX_FUSED_DISABLE=1).  `tools/stress_shapes.py` is the long form of this
test; it found the two-tile ownership bug of wide adjoint bands (n2 = 68, fp32)."""
import os
import random

import pytest
import torch

import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2])
def test_random_shapes_fused_vs_axis_kernels(seed):
    rnd = random.Random(seed)
    for case in range(30):
        dtype = rnd.choice([torch.float32, torch.float64])
        V = 4 if dtype == torch.float32 else 2
        acc = rnd.choice([2, 4, 4, 6])
        lo = 2 * acc + 2
        if rnd.random() < 0.5:
            n0 = rnd.randint(lo, 40)
            n1 = max(lo, rnd.choice([15, 16, 17, 31, 32, 33, 48, 49]))
            n2 = max(V * ((lo + V - 1) // V), V * rnd.choice([8, 15, 16, 17, 18, 31, 32, 33, 34, 49]))
        else:
            n0, n1, n2 = rnd.randint(lo, 60), rnd.randint(lo, 90), V * rnd.randint((lo + V - 1) // V, 200 // V)
        op = rnd.choice(["laplacian", "gradient", "divergence", "curl"])
        h = (rnd.uniform(0.05, 0.5), rnd.uniform(0.05, 0.5), rnd.uniform(0.05, 0.5))
        shape = (n0, n1, n2) if op in ("laplacian", "gradient") else (3, n0, n1, n2)
        x = torch.randn(shape, device="cuda", dtype=dtype)
        kw = dict(accuracy=acc, step_size=h)
        if op == "divergence":
            kw["keepdims"] = False

        def run():
            a = x.clone().requires_grad_(True)
            out = getattr(fdx, op)(a, **kw)
            k = _cabi.last_kernel()
            g = torch.randn(out.shape, device="cuda", dtype=dtype, generator=torch.Generator(device="cuda").manual_seed(case))
            (ga,) = torch.autograd.grad(out, a, g)
            return out.detach(), ga, k

        old = os.environ.get("FDX_FUSED_DISABLE")
        try:
            os.environ["FDX_FUSED_DISABLE"] = "0"
            _cabi.reload_env()  # the library snapshots the FDX_* switches
            out, ga, k = run()
            os.environ["FDX_FUSED_DISABLE"] = "1"
            _cabi.reload_env()
            out_r, ga_r, _ = run()
        finally:
            if old is None:
                os.environ.pop("FDX_FUSED_DISABLE", None)
            else:
                os.environ["FDX_FUSED_DISABLE"] = old
            _cabi.reload_env()
        tol = 3e-5 if dtype == torch.float32 else 2e-11
        e1 = ((out - out_r).abs().max() / out_r.abs().max()).item()
        e2 = ((ga - ga_r).abs().max() / ga_r.abs().max()).item()
        assert e1 <= tol and e2 <= tol, (case, op, dtype, acc, (n0, n1, n2), k, e1, e2)


@pytest.mark.parametrize("seed", [5, 6])
def test_random_operators_vs_oracle(seed):
    """Random small N-D shapes, axes, derivatives, accuracies and methods of every public operator against the
    oracle on the parity metric (axis kernels with their band rows riding along, accumulate paths, scalar
    fallbacks for unaligned shapes, 2-D curl, jacobian / hessian compositions)."""
    import numpy as np

    from oracle import fdx_oracle as O
    from tests._parity import TOL, parity_error

    rnd = random.Random(seed)
    rng = np.random.default_rng(seed)
    done = 0
    for case in range(70):
        dtype = rnd.choice([np.float32, np.float64])
        method = rnd.choice(["central", "central", "forward", "backward"])
        acc = rnd.randint(1, 6)
        op = rnd.choice(["difference", "difference", "laplacian", "gradient", "divergence", "curl", "jacobian", "hessian"])
        big = lambda: rnd.randint(2 * (acc + 3) + 1, 40)  # long enough for every stencil of this accuracy
        if op == "difference":
            nd = rnd.randint(1, 4)
            shape = tuple(big() for _ in range(nd))
            kw = dict(axis=rnd.randrange(-nd, nd), derivative=rnd.randint(1, 3), accuracy=acc, method=method, step_size=rnd.uniform(0.1, 2.0))
        elif op in ("laplacian", "gradient"):
            nd = rnd.randint(1, 3)
            shape = tuple(big() for _ in range(nd))
            kw = dict(accuracy=acc, method=method, step_size=tuple(rnd.uniform(0.1, 2.0) for _ in range(nd)))
        elif op == "divergence":
            nd = rnd.randint(2, 3)
            shape = (nd + rnd.randint(0, 1),) + tuple(big() for _ in range(nd))
            kw = dict(accuracy=acc, method=method, step_size=rnd.uniform(0.1, 2.0), keepdims=rnd.random() < 0.5)
        elif op == "curl":
            nd = rnd.randint(2, 3)
            shape = (nd,) + tuple(big() for _ in range(nd))
            kw = dict(accuracy=acc, method=method, step_size=rnd.uniform(0.1, 2.0))
        elif op == "jacobian":
            nd = rnd.randint(2, 3)
            shape = (rnd.randint(1, 3),) + tuple(big() for _ in range(nd))
            kw = dict(accuracy=acc, method=method, step_size=rnd.uniform(0.1, 2.0))
        else:  # hessian: 2-D (the 3-D key collision of the reference is covered elsewhere)
            shape = (big(), big())
            kw = dict(accuracy=max(acc, 2), method=method, step_size=rnd.uniform(0.1, 2.0))
        x = rng.standard_normal(shape).astype(dtype)
        ref = getattr(O, op)(x, **kw)
        out = getattr(fdx, op)(torch.from_numpy(x).cuda(), **kw).cpu().numpy()
        assert out.shape == ref.shape and out.dtype == ref.dtype, (case, op, shape, kw)
        err = parity_error(op, x, kw, out, ref)
        # both sides use the same exact coefficients (O.* defaults to coeff_mode="exact"): the strict bound holds
        assert err <= TOL[np.dtype(dtype)], (case, op, shape, kw, err)
        done += 1
    assert done == 70


def test_random_shapes_through_the_tma_ring_kernels():
    """tools/stress_tma.py: 200 random `difference` / 2-D operator calls forced through the persistent TMA-ring kernels
    (all methods, derivative 1-4, accuracy 1-8, 2-D to 4-D arrays, every axis) equal the non-persistent kernels bit for
    bit (those are the ones checked against the oracle throughout tests/test_gpu_parity.py)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "stress_tma.py"), "200", "11"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0 and "0 mismatches" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])


def test_fused_hessian3d_random_shapes_against_the_composition():
    """Random 3-D hessians: the one-pass kernel (fdx_hessian3d: chunks, anchored edge tiles, helper-lane x bands, band-row warps,
    z-band CTAs) against the composition of axis kernels (itself checked against the oracle), over shapes from a single
    partial tile to several tiles and chunks, all methods, per-axis accuracies and both dtypes."""
    import numpy as np

    rnd = np.random.default_rng(2024)
    fused = 0
    for case in range(120):
        dtype = torch.float32 if case % 3 else torch.float64
        method = ("central", "forward", "backward")[int(rnd.integers(0, 3))]
        amax = 6 if method == "central" else 3
        acc = tuple(int(a) for a in rnd.integers(1 if method != "central" else 2, amax + 1, size=3))
        if rnd.random() < 0.4:
            acc = acc[0]
        v = 4 if dtype == torch.float32 else 2
        lo = 2 * (amax + 2)
        shape = (int(rnd.integers(lo, 48)), int(rnd.integers(lo, 40)), int(rnd.integers(lo // v + 1, 280 // v)) * v)
        steps = tuple(float(h) for h in rnd.uniform(0.1, 1.0, size=3))
        kw = dict(accuracy=acc, step_size=steps, method=method)
        x = torch.randn(shape, device="cuda", dtype=dtype, generator=torch.Generator(device="cuda").manual_seed(case))
        zc = int(rnd.choice([0, 0, 16, 24]))
        os.environ["FDX_HESS_ZC"] = str(zc)
        _cabi.reload_env()
        try:
            try:
                h = fdx.hessian(x, **kw)
            except ValueError:
                continue  # axis too short for this stencil: the reference's own error
            kern = _cabi.last_kernel()
            os.environ["FDX_HESS_DISABLE"] = "1"
            _cabi.reload_env()
            h0 = fdx.hessian(x, **kw)
            assert _cabi.last_kernel() != "fused_hessian3d"
        finally:
            os.environ.pop("FDX_HESS_DISABLE", None)
            os.environ.pop("FDX_HESS_ZC", None)
            _cabi.reload_env()
        if kern != "fused_hessian3d":
            continue
        fused += 1
        sc = h0.abs().amax(dim=(2, 3, 4), keepdim=True).clamp_min(1e-30)
        err = ((h - h0).abs() / sc).max().item()
        assert err <= (3e-5 if dtype == torch.float32 else 1e-11), (case, shape, kw, zc, err)
    assert fused >= 80, fused


def test_fused_hessian2d_random_shapes_against_the_composition():
    """Random 2-D hessians: the one-pass kernel (fdx_hessian2d: strips, anchored last strip, band columns by shuffles, band
    rows as items of their own, several chunks of rows) against the composition of axis kernels."""
    import numpy as np

    rnd = np.random.default_rng(77)
    fused = 0
    for case in range(120):
        dtype = torch.float32 if case % 3 else torch.float64
        method = ("central", "forward", "backward")[int(rnd.integers(0, 3))]
        amax = 6 if method == "central" else 3
        acc = tuple(int(a) for a in rnd.integers(1 if method != "central" else 2, amax + 1, size=2))
        if rnd.random() < 0.4:
            acc = acc[0]
        v = 4 if dtype == torch.float32 else 2
        lo = 2 * (amax + 2)
        shape = (int(rnd.integers(lo, 200)), int(rnd.integers(lo // v + 1, 600 // v)) * v)
        steps = tuple(float(h) for h in rnd.uniform(0.1, 1.0, size=2))
        kw = dict(accuracy=acc, step_size=steps, method=method)
        x = torch.randn(shape, device="cuda", dtype=dtype, generator=torch.Generator(device="cuda").manual_seed(case))
        os.environ["FDX_HESS2D_WPS"] = str(int(rnd.choice([32, 1, 1000])))
        _cabi.reload_env()
        try:
            try:
                h = fdx.hessian(x, **kw)
            except ValueError:
                continue
            kern = _cabi.last_kernel()
            os.environ["FDX_HESS_DISABLE"] = "1"
            _cabi.reload_env()
            h0 = fdx.hessian(x, **kw)
            assert _cabi.last_kernel() != "fused_hessian2d"
        finally:
            os.environ.pop("FDX_HESS_DISABLE", None)
            os.environ.pop("FDX_HESS2D_WPS", None)
            _cabi.reload_env()
        if kern != "fused_hessian2d":
            continue
        fused += 1
        sc = h0.abs().amax(dim=(2, 3), keepdim=True).clamp_min(1e-30)
        err = ((h - h0).abs() / sc).max().item()
        assert err <= (3e-5 if dtype == torch.float32 else 1e-11), (case, shape, kw, err)
    assert fused >= 80, fused

