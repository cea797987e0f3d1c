"""Axis plans: the host-side description of one 1-D finite-difference operator.

``difference`` along one axis (reference ``finitediffx/_src/finite_diff.py:45-165`` row regions,
``:247-297`` dispatch) is a banded n x n matrix: every *main* row ``i`` applies the same taps at
offsets ``lo..hi``; a few rows next to each end use their own dense taps (the reference's
one-sided boundary stencils).  An ``AxisPlan`` stores exactly that:

    out[i] = sum_k main[k] * x[i + lo + k]                 nl <= i < n - nr
    out[i] = sum_c band_l[i, c] * x[c]                     i < nl          (c < wl)
    out[i] = sum_c band_r[i-(n-nr), c] * x[n - wr + c]     i >= n - nr     (c < wr)

The transposed operator -- what ``jax.grad`` through the reference applies to the cotangent --
has the same shape with wider bands, so ``AxisPlan.transpose()`` is again an ``AxisPlan`` and the
same CUDA kernels run the adjoint ("the transposed stencil, boundary rows included").

Coefficients are float64, solved exactly on the host (``utils._coeffs_f64``) and *unscaled*:
``1/step_size**derivative`` is a runtime scalar handed to the kernel (``step_size`` is a traced
value in the reference, finite_diff.py:168, so it must not be baked into a cached plan).
"""

from __future__ import annotations

import dataclasses as dc
import functools as ft
from typing import Callable, Optional

import numpy as np

from .utils import (
    _coeffs_f64,
    _generate_backward_offsets,
    _generate_central_offsets,
    _generate_forward_offsets,
)

MAX_TAPS = 16  # must match FDX_MAX_TAPS in include/fdx_b200.h

# Hook used by the parity tests to feed the kernels the reference's own (inexact) coefficients,
# SURVEY.md section 9 decision 2(i).  None -> exact float64.
_coeff_override: Optional[Callable] = None
_dependent_caches: list = []  # clear() callables of caches that hold plans (finite_diff._fused_memo ...)


def set_coefficient_source(fn: Optional[Callable]) -> None:
    """``fn(offsets, derivative) -> float64 array`` replaces the exact solve (tests only)."""
    global _coeff_override
    _coeff_override = fn
    axis_plan.cache_clear()
    axis_plan_t.cache_clear()
    for clear in _dependent_caches:
        clear()


def _coeffs(offsets, derivative):
    if _coeff_override is not None:
        return np.asarray(_coeff_override(tuple(offsets), derivative), dtype=np.float64)
    return _coeffs_f64(tuple(offsets), derivative)


@dc.dataclass(frozen=True, eq=False)
class AxisPlan:
    n: int
    lo: int
    hi: int
    main: np.ndarray  # (hi-lo+1,) float64
    band_l: np.ndarray  # (nl, wl) float64
    band_r: np.ndarray  # (nr, wr) float64
    key: tuple = ()

    @property
    def nl(self) -> int:
        return self.band_l.shape[0]

    @property
    def wl(self) -> int:
        return self.band_l.shape[1]

    @property
    def nr(self) -> int:
        return self.band_r.shape[0]

    @property
    def wr(self) -> int:
        return self.band_r.shape[1]

    @property
    def taps(self) -> int:
        return self.hi - self.lo + 1

    def resized(self, n: int) -> "AxisPlan":
        """Same operator family on an axis of another length (bands stay glued to the ends)."""
        assert n >= self.nl + self.nr and n >= max(self.wl, self.wr)
        return dc.replace(self, n=n)

    def dense(self, n: Optional[int] = None) -> np.ndarray:
        """The full n x n matrix (tests, small-n fallbacks and transposition)."""
        n = self.n if n is None else n
        assert n >= self.nl + self.nr and n >= max(self.wl, self.wr)
        m = np.zeros((n, n), dtype=np.float64)
        m[: self.nl, : self.wl] = self.band_l
        if self.nr:
            m[n - self.nr :, n - self.wr :] = self.band_r
        for i in range(self.nl, n - self.nr):
            m[i, i + self.lo : i + self.hi + 1] = self.main
        return m

    def transpose(self) -> "AxisPlan":
        """Plan of the transposed matrix."""
        n, lo, hi, nl, nr, wl, wr = self.n, self.lo, self.hi, self.nl, self.nr, self.wl, self.wr
        if nl + nr >= n:  # already dense
            return _dense_plan(self.dense().T.copy(), key=("T",) + self.key)
        # rows of D^T that are NOT the plain reflected stencil: columns of D touched by a band
        # row, or by fewer than all `taps` main rows
        tnl = max(wl, nl + hi)
        tnr = max(wr, nr - lo)
        twl = max(nl, tnl - lo)
        twr = max(nr, tnr + hi)
        if tnl + tnr > n or twl > n or twr > n:
            return _dense_plan(self.dense().T.copy(), key=("T",) + self.key)
        # the middle of the matrix is translation invariant, so the corners of a small
        # representative instance are the corners of the real one
        nrep = min(n, tnl + tnr + 2 * (hi - lo + 1) + 4 + twl + twr)
        t = self.dense(nrep).T
        return AxisPlan(
            n=n,
            lo=-hi,
            hi=-lo,
            main=self.main[::-1].copy(),
            band_l=np.ascontiguousarray(t[:tnl, :twl]),
            band_r=np.ascontiguousarray(t[nrep - tnr :, nrep - twr :]),
            key=("T",) + self.key,
        )


def _dense_plan(matrix: np.ndarray, key=()) -> AxisPlan:
    n = matrix.shape[0]
    return AxisPlan(
        n=n,
        lo=0,
        hi=0,
        main=np.zeros(1),
        band_l=np.ascontiguousarray(matrix, dtype=np.float64),
        band_r=np.zeros((0, 0)),
        key=key,
    )


def _too_short(axis_label, accuracy, derivative, need, size, method):
    # same wording contract as finite_diff.py:292-297 (ValueError for short axis / bad method)
    return ValueError(
        f"Size of the array along axis {axis_label} is smaller than the number of points required"
        f" for the accuracy {accuracy} and derivative {derivative} requested"
        f".\nSize must be larger than {need}, but got {size}"
        f".\nMethod should be 'central', 'forward', 'backward', but got {method}"
    )


@ft.lru_cache(maxsize=1024)
def axis_plan(n: int, derivative: int, accuracy: int, method: str, axis_label: int = 0) -> AxisPlan:
    """Forward plan of ``difference`` along an axis of length ``n``.

    Follows the dispatch of finite_diff.py:247-297 and the row regions of :45-165 (see the row map
    in SURVEY.md section 3.2).  Sizes for which the reference's own slices leave the array (it
    then dies inside ``lax.slice`` instead of raising its ``ValueError``) raise ``ValueError`` here.
    """
    d, a = int(derivative), int(accuracy)
    cen_off = _generate_central_offsets(d, a + 1)
    fwd_off = _generate_forward_offsets(d, a)
    bwd_off = _generate_backward_offsets(d, a)
    cen, fwd, bwd = _coeffs(cen_off, d), _coeffs(fwd_off, d), _coeffs(bwd_off, d)
    p, L = cen_off[-1], fwd_off[-1]
    key = (d, a, method)

    def left_band(rows):  # rows [0, rows) use the forward stencil
        b = np.zeros((rows, rows + L))
        for i in range(rows):
            b[i, i : i + L + 1] = fwd
        return b

    def right_band(rows):  # rows [n-rows, n) use the backward stencil
        b = np.zeros((rows, rows + L))
        for i in range(rows):
            b[i, i : i + L + 1] = bwd
        return b

    if method == "central" and n >= len(cen_off):
        if n < p + L:
            raise _too_short(axis_label, a, d, p + L - 1, n, method)
        return AxisPlan(n, -p, p, cen, left_band(p), right_band(p), key)
    if method == "central" and n >= len(fwd_off):
        # _central_difference_edge (:93-115): only n == d+a reaches here; the slices are in
        # bounds only when 2L + n%2 <= n, i.e. the two-point case d+a == 2.
        rows_l, rows_r = L + n % 2, L
        if rows_l + rows_r != n or 2 * L + n % 2 > n:
            raise _too_short(axis_label, a, d, p + L - 1, n, method)
        m = np.zeros((n, n))
        for i in range(rows_l):
            m[i, i : i + L + 1] = fwd
        for i in range(rows_r):
            r = n - rows_r + i
            m[r, r - L : r + 1] = bwd
        return _dense_plan(m, key)
    if method == "forward" and n >= len(fwd_off):
        if n < 2 * L:
            raise _too_short(axis_label, a, d, 2 * L - 1, n, method)
        return AxisPlan(n, 0, L, fwd, np.zeros((0, 0)), right_band(L), key)
    if method == "backward" and n >= len(bwd_off):
        if n < 2 * L:
            raise _too_short(axis_label, a, d, 2 * L - 1, n, method)
        return AxisPlan(n, -L, 0, bwd, left_band(L), np.zeros((0, 0)), key)
    raise _too_short(axis_label, a, d, len(fwd_off), n, method)


@ft.lru_cache(maxsize=1024)
def axis_plan_t(n: int, derivative: int, accuracy: int, method: str) -> AxisPlan:
    """Transposed (adjoint) plan."""
    return axis_plan(n, derivative, accuracy, method).transpose()
