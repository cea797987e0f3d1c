"""Multi-threaded CPU arm for bench.py (``--impl reference`` and ``cpu_baseline``).  TEST/BENCH
INFRASTRUCTURE ONLY -- never imported by the product.

The reference's jax[cpu] path cannot run here (jax is not installed in this image).
This restates the same graph the reference traces -- per region: slice -> scale -> sequential
sum; then concatenate -> divide; then the cross-axis sum (finitediffx/_src/finite_diff.py:45-165,
234-297, 443-453, 565-575) -- on torch CPU tensors, so the elementwise work runs on all host cores
the way XLA:CPU's thread pool would run it.  Label every number from here
"restated reference algorithm, torch-CPU, N cores -- not jax[cpu]".

It is checked against the NumPy oracle in tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np
import torch

from . import fdx_oracle as O


def _difference(x: torch.Tensor, *, axis, accuracy, step_size, derivative, method):
    n = x.shape[axis]
    dt = np.float64 if x.dtype == torch.float64 else np.float32
    cen, fwd, bwd = O.stencil_set(derivative, accuracy, "exact", dt)
    kind = O._dispatch(method, n, cen, fwd, bwd)
    parts = []
    for terms in O._region_terms(kind, n, cen, fwd, bwd):
        total = None
        for coeff, (start, stop) in terms:
            if not (0 <= start <= stop <= n):
                raise ValueError("slice outside the axis")
            piece = x.narrow(axis, start, stop - start) * float(coeff)
            total = piece if total is None else total + piece
        parts.append(total)
    return torch.cat(parts, dim=axis) / (float(step_size) ** derivative)


def laplacian(x, *, accuracy=1, step_size=1, method="central"):
    acc = O._per_axis(accuracy, x.ndim, "accuracy")
    step = O._per_axis(step_size, x.ndim, "step_size")
    total = None
    for k, (a, h) in enumerate(zip(acc, step)):
        term = _difference(x, axis=k, accuracy=a, step_size=h, derivative=2, method=method)
        total = term if total is None else total + term
    return total


def divergence(x, *, accuracy=1, step_size=1, keepdims=True, method="central"):
    acc = O._per_axis(accuracy, x.ndim - 1, "accuracy")
    step = O._per_axis(step_size, x.ndim - 1, "step_size")
    total = None
    for k, (a, h) in enumerate(zip(acc, step)):
        term = _difference(x[k], axis=k, accuracy=a, step_size=h, derivative=1, method=method)
        total = term if total is None else total + term
    return total[None] if keepdims else total


def gradient(x, *, accuracy=1, step_size=1, method="central"):
    acc = O._per_axis(accuracy, x.ndim, "accuracy")
    step = O._per_axis(step_size, x.ndim, "step_size")
    return torch.stack([_difference(x, axis=k, accuracy=a, step_size=h, derivative=1, method=method) for k, (a, h) in enumerate(zip(acc, step))])


def difference(x, **kw):
    return _difference(x, **kw)


def curl(x, *, accuracy=1, step_size=1, method="central", keepdims=True):
    """3-D curl as the reference composes it (finitediffx/_src/finite_diff.py:616-668): six differences, three
    subtractions, one stack."""
    if not (x.ndim == 4 and x.shape[0] == 3):
        raise ValueError("the CPU arm times the 3-D curl only")
    acc = O._per_axis(accuracy, 3, "accuracy")
    step = O._per_axis(step_size, 3, "step_size")
    d = lambda c, ax: _difference(x[c], axis=ax, accuracy=acc[ax], step_size=step[ax], derivative=1, method=method)
    return torch.stack([d(2, 1) - d(1, 2), d(0, 2) - d(2, 0), d(1, 0) - d(0, 1)])
