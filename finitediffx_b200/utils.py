"""Offsets, finite-difference coefficients and argument broadcasting (host side).

Mirrors the reference's ``finitediffx/_src/utils.py``:

* ``_generate_forward_offsets`` / ``_generate_central_offsets`` / ``_generate_backward_offsets``
  follow ``utils.py:36-54`` / ``57-77`` / ``80-97`` (same ranges, same ``ValueError`` rules).
* ``generate_finitediff_coeffs`` follows ``utils.py:114-145`` -> ``100-111``: it solves
  ``sum_j c_j * o_j**i = derivative! * [i == derivative]``.  The reference inverts the
  Vandermonde matrix on the device at the ambient JAX width (int32/float32 unless x64 is on);
  here the system is solved on the host in exact rational arithmetic and rounded once to
  float64 (north star: "solved on the host in float64 and passed to the kernels as constants").
* ``_check_and_return`` follows ``utils.py:25-33``.

Nothing in this module touches the GPU.
"""

from __future__ import annotations

import functools as ft
import math
from fractions import Fraction
from typing import Any, Sequence

import numpy as np

__all__ = (
    "generate_finitediff_coeffs",
    "_generate_forward_offsets",
    "_generate_central_offsets",
    "_generate_backward_offsets",
    "_check_and_return",
)


def _generate_forward_offsets(derivative: int, accuracy: int = 2) -> tuple[int, ...]:
    """Forward offsets ``0 .. derivative+accuracy-1`` (reference ``utils.py:36-54``)."""
    if derivative < 1:
        raise ValueError(f"{derivative=} <1 for forward offset generation")
    if accuracy < 1:
        raise ValueError(f"{accuracy=} <1 for forward offset generation")
    return tuple(range(0, derivative + accuracy))


def _generate_central_offsets(derivative: int, accuracy: int = 2) -> tuple[int, ...]:
    """Central offsets ``-q .. q`` with ``q=(derivative+accuracy-1)//2`` (``utils.py:57-77``)."""
    if derivative < 1:
        raise ValueError(f"{derivative=} <1 for central offset generation")
    if accuracy < 2:
        raise ValueError(f"{accuracy=} <2 for central offset generation")
    half = (derivative + accuracy - 1) // 2
    return tuple(range(-half, half + 1))


def _generate_backward_offsets(derivative: int, accuracy: int = 2) -> tuple[int, ...]:
    """Backward offsets ``-(derivative+accuracy-1) .. 0`` (reference ``utils.py:80-97``)."""
    if derivative < 1:
        raise ValueError(f"{derivative=} <1 for back offset generation")
    if accuracy < 1:
        raise ValueError(f"{accuracy=} <1 for back offset generation")
    return tuple(range(-(derivative + accuracy - 1), 1))


def _to_fraction(value) -> Fraction:
    # Fraction(float) is exact for the binary value, so non-integer offsets such as -2.2
    # (reference tests/test_fgrad.py:37) are handled without any rounding of our own.
    if isinstance(value, (int, np.integer)):
        return Fraction(int(value))
    return Fraction(float(value))


@ft.lru_cache(maxsize=4096)
def _solve_exact(offsets: tuple, derivative: int) -> tuple[Fraction, ...]:
    """Exact rational solution of the Vandermonde system of ``utils.py:104-111``."""
    n = len(offsets)
    offs = [_to_fraction(o) for o in offsets]
    # augmented matrix [A | B], A[i][j] = offs[j]**i, B[i] = derivative! * (i == derivative)
    rows = [[o**i for o in offs] + [Fraction(math.factorial(derivative) if i == derivative else 0)] for i in range(n)]
    for col in range(n):
        pivot = next((r for r in range(col, n) if rows[r][col] != 0), None)
        if pivot is None:
            raise ValueError(f"offsets={offsets} are not distinct; the stencil system is singular")
        rows[col], rows[pivot] = rows[pivot], rows[col]
        inv = 1 / rows[col][col]
        rows[col] = [v * inv for v in rows[col]]
        for r in range(n):
            if r != col and rows[r][col] != 0:
                f = rows[r][col]
                rows[r] = [a - f * b for a, b in zip(rows[r], rows[col])]
    return tuple(rows[i][n] for i in range(n))


def _coeffs_f64(offsets: Sequence, derivative: int) -> np.ndarray:
    """float64 coefficients, each the correctly rounded value of the exact rational."""
    exact = _solve_exact(tuple(offsets), int(derivative))
    return np.array([float(c) for c in exact], dtype=np.float64)


def generate_finitediff_coeffs(offsets: Sequence[float | int], derivative: int) -> np.ndarray:
    """Generate FD coeffs (drop-in for the reference's ``generate_finitediff_coeffs``).

    Args:
        offsets: offsets of the finite difference stencil
        derivative: derivative order

    Returns:
        float64 ``numpy.ndarray`` of finite difference coefficients (the reference returns a
        ``jax.Array`` at the ambient width; see DESIGN.md "coefficient provenance").

    Example:
        >>> generate_finitediff_coeffs(offsets=(-1, 0, 1), derivative=1)
        array([-0.5,  0. ,  0.5])
        >>> generate_finitediff_coeffs(offsets=(-1, 0, 1), derivative=2)
        array([ 1., -2.,  1.])
    """
    if derivative >= len(offsets):
        raise ValueError(f"{offsets=} of {len(offsets)=} must be larger than {derivative=}.")
    return _coeffs_f64(tuple(offsets), derivative)


def _is_scalar_array(value: Any) -> bool:
    # torch tensors / numpy arrays / numpy scalars play the role of ``jax.Array`` here
    try:
        import torch

        if isinstance(value, torch.Tensor):
            return True
    except Exception:  # pragma: no cover
        pass
    return isinstance(value, (np.ndarray, np.generic))


def _check_and_return(value: Any, ndim: int, name: str):
    """Broadcast ``accuracy`` / ``step_size`` to one entry per axis (reference ``utils.py:25-33``).

    int -> repeated tuple; array (the ``jax.Array`` branch) -> repeated; tuple -> length checked
    with ``assert`` exactly like the reference; anything else -> ``ValueError``.  A Python float
    is accepted for ``step_size`` only: in the reference it reaches this function as a tracer
    under the outer ``jax.jit`` and takes the ``jax.Array`` branch, whereas ``accuracy`` is a
    static argument and stays a Python object.
    """
    if isinstance(value, bool):
        raise ValueError(f"Expected int or tuple for {name}, got {value}.")
    if isinstance(value, int):
        return (value,) * ndim
    if _is_scalar_array(value):
        flat = np.asarray(value.detach().cpu() if hasattr(value, "detach") else value).reshape(-1)
        if flat.size == 1:
            return (flat[0].item(),) * ndim
        # jnp.repeat(value, ndim) on a length-k array gives k*ndim entries, which the
        # reference then zips against the axes; keep that truncating behaviour.
        return tuple(np.repeat(flat, ndim).tolist())
    if isinstance(value, float) and name == "step_size":
        return (value,) * ndim
    if isinstance(value, tuple):
        assert len(value) == ndim, f"{name} must be a tuple of length {ndim}"
        return tuple(value)
    raise ValueError(f"Expected int or tuple for {name}, got {value}.")
