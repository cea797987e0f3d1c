"""finitediffx_b200 -- B200-native array finite-difference operators.

Drop-in for the array half of ASEM000/finitediffX (``finitediffx/__init__.py:15-40``): the same
eight public names with unchanged signatures, computed by hand-written CUDA (sm_100a) behind the
C ABI in ``include/fdx_b200.h``.  ``fgrad`` / ``value_and_fgrad`` / ``define_fdjvp`` / ``Offset``
(black-box function differentiation, ``finitediffx/_src/fgrad.py``) are out of scope: they never
touch the array stencil path and stay the reference's Python.
"""

from .finite_diff import (
    curl,
    difference,
    divergence,
    gradient,
    hessian,
    jacobian,
    laplacian,
)
from .utils import generate_finitediff_coeffs

__all__ = (
    "curl",
    "divergence",
    "difference",
    "gradient",
    "jacobian",
    "laplacian",
    "hessian",
    "generate_finitediff_coeffs",
)

__version__ = "0.1.0"
