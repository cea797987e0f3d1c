"""Array finite-difference operators: drop-in for ``finitediffx/_src/finite_diff.py``.

Same names, keyword-only signatures, defaults and error behaviour as the reference
(``difference`` :168-297, ``gradient`` :300-351, ``jacobian`` :354-401, ``divergence`` :404-457,
``hessian`` :460-529, ``laplacian`` :532-575, ``curl`` :671-741); the arithmetic runs in the
hand-written CUDA library behind include/fdx_b200.h.

Every operator is a short list of *terms* ``out[co] += alpha * D(axis, derivative, accuracy,
method) in[ci]``; ``_StencilOp`` executes a term list, picking the fused single-pass 3-D kernels
when the list is a laplacian / divergence / gradient / curl on a 3-D grid and the generic axis
kernels (accumulating in place, no temporaries, no stack copy) otherwise.  Its backward pass
executes the transposed term list with transposed plans -- boundary rows included -- which is what
``jax.grad`` through the reference computes; it is the
``torch.autograd.Function`` analogue of the ``jax.custom_vjp`` the north star asks for.

Arrays: CUDA ``torch.Tensor`` in -> CUDA tensor out.  Host data (NumPy arrays, lists, CPU tensors)
is copied to the current CUDA device, processed there, and returned in the kind it came in.
"""

from __future__ import annotations

import dataclasses as dc
from itertools import combinations
from typing import Literal, Sequence

import numpy as np
import torch

from . import _cabi
from . import plan as _plan_mod
from .plan import AxisPlan, axis_plan, axis_plan_t
from .utils import _check_and_return

__all__ = ("hessian", "difference", "gradient", "jacobian", "laplacian", "curl", "divergence")

MethodKind = Literal["central", "forward", "backward"]


# ------------------------------------------------------------------------------- term lists
@dc.dataclass(frozen=True)
class Term:
    co: int  # output field index
    ci: int  # input field index
    axis: int  # spatial axis
    derivative: int
    accuracy: int
    method: str
    alpha: float  # runtime scale, 1 / step**derivative (with sign)
    transposed: bool = False

    def plan(self, n: int) -> AxisPlan:
        if self.transposed:
            return axis_plan_t(n, self.derivative, self.accuracy, self.method)
        return axis_plan(n, self.derivative, self.accuracy, self.method, self.axis)

    def t(self) -> "Term":
        return dc.replace(self, co=self.ci, ci=self.co, transposed=not self.transposed)


def _match_fused(terms: Sequence[Term], n_in: int, n_out: int, spatial):
    """Recognise the four fused 3-D shapes.  Returns (kind, per-axis terms) or None."""
    if len(spatial) != 3 or not terms:
        return None
    # every axis brings its own plan (derivative, accuracy and method may differ from axis to axis: the kernels run
    # at the largest radius and pad shorter / one-sided stencils with zero taps); only the adjoint flag is shared
    t0 = terms[0]
    if any(t.transposed != t0.transposed for t in terms):
        return None
    by = {(t.co, t.ci, t.axis): t for t in terms}
    if len(by) != len(terms):
        return None
    if len(terms) == 3:
        if n_in == 1 and n_out == 1 and set(by) == {(0, 0, k) for k in range(3)}:
            return "laplacian", [by[(0, 0, k)] for k in range(3)]
        if n_out == 1 and n_in >= 3 and set(by) == {(0, k, k) for k in range(3)}:
            return "divergence", [by[(0, k, k)] for k in range(3)]
        if n_in == 1 and n_out >= 3 and set(by) == {(k, 0, k) for k in range(3)}:
            return "gradient", [by[(k, 0, k)] for k in range(3)]
    if len(terms) == 6 and n_in == 3 and n_out == 3:
        # out_0 = a1 D1 x2 - a2 D2 x1 ; out_1 = a2 D2 x0 - a0 D0 x2 ; out_2 = a0 D0 x1 - a1 D1 x0
        plus = {(0, 2, 1), (1, 0, 2), (2, 1, 0)}
        minus = {(0, 1, 2), (1, 2, 0), (2, 0, 1)}
        if set(by) == plus | minus:
            axis_terms = []
            for k in range(3):
                tp = next(by[key] for key in plus if key[2] == k)
                tm = next(by[key] for key in minus if key[2] == k)
                if tp.alpha != -tm.alpha or (tp.derivative, tp.accuracy, tp.method) != (tm.derivative, tm.accuracy, tm.method):
                    return None
                axis_terms.append(tp)
            return "curl", axis_terms
    return None


_fused_memo: dict = {}
_plan_mod._dependent_caches.append(_fused_memo.clear)


def _match_groups(terms: Sequence[Term], n_in: int, n_out: int, spatial):
    """A term list as fused launches: [(kind, axis terms, first input field, first output field)] or None.
    Either the whole list is one fused operator, or it splits into independent gradients of single input fields with
    consecutive output fields -- the jacobian (finite_diff.py:392-401: one gradient per component, C x D axis passes
    in the reference) becomes C single-pass launches -- or, transposed, into divergences (the jacobian's vjp)."""
    m = _match_fused(terms, n_in, n_out, spatial)
    if m is not None:
        return [(m[0], m[1], 0, 0)]
    if len(spatial) != 3 or not terms or len(terms) % 3:
        return None
    groups = []
    for by_in in (True, False):  # gradients grouped by input field / divergences grouped by output field
        buckets: dict = {}
        for t in terms:
            buckets.setdefault(t.ci if by_in else t.co, []).append(t)
        groups = []
        for key, ts in sorted(buckets.items()):
            base = min((t.co if by_in else t.ci) for t in ts)
            local = [dc.replace(t, co=t.co - base, ci=0) if by_in else dc.replace(t, ci=t.ci - base, co=0) for t in ts]
            m = _match_fused(tuple(local), 1 if by_in else 3, 3 if by_in else 1, spatial)
            if m is None or m[0] != ("gradient" if by_in else "divergence"):
                groups = None
                break
            groups.append((m[0], m[1], key if by_in else base, base if by_in else key))
        if groups:
            return groups
    return None


def _match_2d(terms: Sequence[Term]):
    """Two terms on a 2-D grid, one per axis, as one fused 2-D launch: (kind, (t_axis0, t_axis1), input fields,
    output fields) or None.  Same output and input: laplacian; same output, two inputs: divergence (the 2-D curl is one
    with a negated scale); two outputs, same input: gradient (and the adjoints of the former two)."""
    if len(terms) != 2 or {t.axis for t in terms} != {0, 1} or terms[0].transposed != terms[1].transposed:
        return None
    t0, t1 = sorted(terms, key=lambda t: t.axis)
    if t0.co == t1.co:
        if t0.ci == t1.ci:
            return "laplacian", (t0, t1), (t0.ci,), (t0.co,)
        return "divergence", (t0, t1), (t0.ci, t1.ci), (t0.co,)
    if t0.ci == t1.ci:
        return "gradient", (t0, t1), (t0.ci,), (t0.co, t1.co)
    return None


def _fused_lookup_2d(terms, n_in: int, n_out: int, spatial, dtype, device):
    key = (id(terms), n_in, n_out, spatial, dtype, device.index)
    hit = _fused_memo.get(key)
    if hit is not None and hit[0] is terms:
        return hit[1]
    vkey = (terms, n_in, n_out, spatial, dtype, device.index, "2d")
    hit = _fused_memo.get(vkey)
    if hit is not None:
        launcher = hit[1]
    else:
        # the whole list, or independent pairs grouped by input field (jacobian: one gradient per component) or by
        # output field (its adjoint: one divergence per component)
        groups = [terms] if len(terms) == 2 else []
        if not groups and len(terms) % 2 == 0:
            for attr in ("ci", "co"):
                buckets: dict = {}
                for t in terms:
                    buckets.setdefault(getattr(t, attr), []).append(t)
                if all(len(b) == 2 for b in buckets.values()):
                    groups = [tuple(b) for b in buckets.values()]
                    break
        matches = [_match_2d(g) for g in groups]
        launcher = None
        if matches and all(m is not None for m in matches):
            launcher = tuple(
                _cabi.FusedLauncher2D(kind, (t0.plan(spatial[0]), t1.plan(spatial[1])), (t0.alpha, t1.alpha), fin, fout, dtype, device)
                for kind, (t0, t1), fin, fout in matches)
        if len(_fused_memo) > 512:
            _fused_memo.clear()
        _fused_memo[vkey] = (terms, launcher)
    _fused_memo[key] = (terms, launcher)
    return launcher


def _fused_lookup(terms, n_in: int, n_out: int, spatial, dtype, device):
    """Memoised recognition of a term list as fused 3-D launches: a tuple of (prepared ``_cabi.FusedLauncher``, first
    input field, first output field) or None.  The key uses the identity of the term tuple (``_memo_terms`` hands out
    the same tuple for repeated calls, and the entry keeps it alive), so a hit costs one small-tuple hash instead of
    hashing every Term."""
    key = (id(terms), n_in, n_out, spatial, dtype, device.index)
    hit = _fused_memo.get(key)
    if hit is not None and hit[0] is terms:
        return hit[1]
    vkey = (terms, n_in, n_out, spatial, dtype, device.index)  # an equal term list built anew (tensor arguments)
    hit = _fused_memo.get(vkey)
    if hit is not None:
        launches = hit[1]
    else:
        groups = _match_groups(terms, n_in, n_out, spatial)
        launches = None
        if groups is not None:
            launches = tuple(
                (_cabi.FusedLauncher(kind, tuple(t.plan(spatial[k]) for k, t in enumerate(axis_terms)), tuple(t.alpha for t in axis_terms), dtype, device), fi, fo)
                for kind, axis_terms, fi, fo in groups)
        if len(_fused_memo) > 512:
            _fused_memo.clear()
        _fused_memo[vkey] = (terms, launches)
    _fused_memo[key] = (terms, launches)
    return launches


def _execute(terms: Sequence[Term], x: torch.Tensor, n_out: int) -> torch.Tensor:
    """x: (n_in, *spatial) contiguous CUDA tensor -> (n_out, *spatial)."""
    n_in, spatial = x.shape[0], tuple(x.shape[1:])
    out = torch.empty((n_out,) + spatial, dtype=x.dtype, device=x.device)
    if x.numel() == 0:
        return out
    launches = _fused_lookup(terms, n_in, n_out, spatial, x.dtype, x.device) if len(spatial) == 3 else None
    if len(spatial) == 2:
        l2 = _fused_lookup_2d(terms, n_in, n_out, spatial, x.dtype, x.device)
        if l2 is not None:
            esz = x.element_size()
            rc, done = 0, set()
            with _cabi._on_device(x.device):
                for one in l2:
                    rc = one.launch(x.data_ptr(), x.stride(0) * esz, out.data_ptr(), out.stride(0) * esz)
                    if rc != 0:
                        break  # (not covered: everything goes through the axis kernels below)
                    done.update(one.fout)
            if rc == 0:
                for co in range(n_out):
                    if co not in done:
                        out[co].zero_()
                return out
    if launches is not None:
        esz = x.element_size()
        xs, os_ = x.stride(0) * esz, out.stride(0) * esz
        xp, op = x.data_ptr(), out.data_ptr()
        ok, covered = True, 0
        with _cabi._on_device(x.device):
            for launcher, fi, fo in launches:
                if launcher.launch(xp + fi * xs, xs, op + fo * os_, os_) != 0:
                    ok = False  # (the fused kernels do not cover this call: everything goes through the axis kernels)
                    break
                covered += launcher.nout
        if ok:
            if covered < n_out:  # fields no term writes (the vjp of a divergence that ignores extra components)
                written = set()
                for launcher, fi, fo in launches:
                    written.update(range(fo, fo + launcher.nout))
                for co in range(n_out):
                    if co not in written:
                        out[co].zero_()
            return out
    written = [False] * n_out
    for t in terms:
        _cabi.axis_apply(t.plan(spatial[t.axis]), x[t.ci], out[t.co], t.axis, t.alpha, accumulate=written[t.co])
        written[t.co] = True
    for co, w in enumerate(written):
        if not w:
            out[co].zero_()
    return out


_transposed_memo: dict = {}
_plan_mod._dependent_caches.append(_transposed_memo.clear)


def _transposed(terms: tuple) -> tuple:
    """The transposed term list, one tuple object per forward list (so the launch caches hit on its identity)."""
    hit = _transposed_memo.get(id(terms))
    if hit is None or hit[0] is not terms:
        if len(_transposed_memo) > 256:
            _transposed_memo.clear()
        hit = _transposed_memo[id(terms)] = (terms, tuple(t.t() for t in terms))
    return hit[1]


class _StencilOp(torch.autograd.Function):
    """Linear stencil operator given by a term list; backward = transposed term list."""

    @staticmethod
    def forward(ctx, x, terms, n_out):
        ctx.terms, ctx.n_in = terms, x.shape[0]
        return _execute(terms, x, n_out)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        return _StencilOp.apply(g, _transposed(ctx.terms), ctx.n_in), None, None


# ------------------------------------------------------------------------ array marshalling
def _identity(r):
    return r


def _to_device(array):
    """-> (contiguous CUDA tensor of float32/float64, restore(result) -> caller's array kind)."""
    if type(array) is torch.Tensor and array.is_cuda and (array.dtype is torch.float32 or array.dtype is torch.float64) and array.is_contiguous():
        _cabi.load()
        return array, _identity  # the common case: nothing to convert, nothing to restore
    if isinstance(array, torch.Tensor):
        t, kind = array, ("cuda" if array.is_cuda else "cpu")
    else:
        a = np.asarray(array)
        if a.dtype == object:
            raise TypeError(f"cannot interpret {type(array)} as a numeric array")
        t, kind = torch.from_numpy(np.ascontiguousarray(a)), "numpy"
    if not torch.cuda.is_available():
        raise _cabi.FdxError("finitediffx_b200 needs a CUDA device (B200); there is no CPU fallback")
    _cabi.load()
    # output dtype rule (reference: promote(coefficients, input)): float64 stays float64,
    # everything else (float32, half, bfloat16, integers, bool) computes in float32
    target = torch.float64 if t.dtype == torch.float64 else torch.float32
    pinned = (not t.is_cuda) and t.is_pinned()
    if not t.is_cuda:
        t = t.to("cuda", non_blocking=pinned)
    t = t.to(target).contiguous()

    def restore(r: torch.Tensor):
        if kind == "cuda":
            return r
        host = torch.empty(r.shape, dtype=r.dtype, pin_memory=True)  # torch caches pinned blocks
        host.copy_(r, non_blocking=True)
        torch.cuda.current_stream(r.device).synchronize()
        return host if kind == "cpu" else host.numpy()

    return t, restore


# ------------------------------------------------------------------ host arrays: streamed slabs
_streamers: dict = {}
_streamer_lock = __import__("threading").Lock()


def _host_streamed(kind: str, array, accuracy, step_size, method: str, n_out: int):
    """Large host-resident 3-D grids: pipeline slabs through the GPU (H2D, kernel, D2H overlapped,
    finitediffx_b200/slab.py) instead of staging the whole array in HBM.  Returns None when the
    call is not of that kind, so that the caller takes the ordinary path."""
    if isinstance(array, torch.Tensor):
        if array.is_cuda or array.requires_grad:
            return None
        t, as_numpy = array, False
    elif isinstance(array, np.ndarray):
        t, as_numpy = torch.from_numpy(np.ascontiguousarray(array)), True
    else:
        return None
    from . import slab as _slab

    nin = _slab.FIELDS[kind][0]
    if t.dtype not in (torch.float32, torch.float64) or not torch.cuda.is_available():
        return None
    if t.ndim != (3 if nin == 1 else 4) or (nin == 3 and t.shape[0] != 3):
        return None
    spatial = tuple(t.shape[-3:])
    if t.numel() < (1 << 24):  # small grids: one copy each way is simpler and as fast
        return None
    acc = _check_and_return(accuracy, 3, "accuracy")
    steps = _check_and_return(step_size, 3, "step_size")
    if len(set(acc)) != 1 or method != "central":
        return None
    dev = torch.device("cuda", torch.cuda.current_device())
    try:
        spec = _slab.make_spec(kind, spatial, accuracy=acc[0], step_size=tuple(steps), method=method)
    except ValueError:
        return None
    key = (kind, spatial, t.dtype, acc[0], tuple(float(h) for h in steps), dev.index, _slab.DEFAULT_SLAB_BYTES)
    st = _streamers.get(key)
    if st is None:
        # a streamer keeps 3 rotating sets of input and output slabs in HBM (~0.4-0.8 GB): at most two stay cached
        while len(_streamers) >= 2:
            _streamers.pop(next(iter(_streamers)))
        st = _streamers[key] = _slab.HostStreamer(spec, t.dtype, dev, slab_planes=max(8, _slab.DEFAULT_SLAB_BYTES // max(1, spatial[1] * spatial[2] * t.element_size())))
    t = t.contiguous()
    if not t.is_pinned():
        pinned = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        pinned.copy_(t)
        t = pinned
    out = torch.empty(((n_out,) if n_out > 1 else ()) + spatial, dtype=t.dtype, pin_memory=True)
    xs = [t] if nin == 1 else [t[0], t[1], t[2]]
    outs = [out] if n_out == 1 else [out[k] for k in range(n_out)]
    try:
        with _streamer_lock:  # the streamer's buffers, streams and events are shared state: one call at a time
            st.run(xs, outs)
    except _cabi.FdxError:
        return None
    torch.cuda.current_stream(dev).synchronize()
    return out.numpy() if as_numpy else out


def _alpha(step, derivative: int) -> float:
    if isinstance(step, torch.Tensor):
        step = step.item()
    return 1.0 / (float(step) ** derivative)


# ------------------------------------------------------------------ traced step_size (a device value)
def _traced_steps(step_size, nd: int, device):
    """``step_size`` as the reference sees it under ``jax.jit`` -- a traced DEVICE value (finite_diff.py:168: it is not a
    static argument): a CUDA tensor, or a tensor that wants a gradient (or a tuple holding one).  Returns a float64
    CUDA vector of ``nd`` step sizes that keeps the autograd graph, or None for concrete Python / NumPy values.  The
    host never reads the values (no ``.item()``, no synchronisation: the call can be captured in a CUDA graph); the
    kernels read the scales 1 / h**d on the device (the *_dev_alpha entry points)."""
    def traced(v):
        return isinstance(v, torch.Tensor) and (v.is_cuda or v.requires_grad)

    if traced(step_size):
        h = step_size.to(device=device, dtype=torch.float64).reshape(-1)
        if h.numel() == 1:
            return h.expand(nd)
        return h.repeat_interleave(nd)[:nd]  # jnp.repeat(value, ndim), zipped against the axes (utils.py:29)
    if isinstance(step_size, tuple) and any(traced(v) for v in step_size):
        assert len(step_size) == nd, f"step_size must be a tuple of length {nd}"
        return torch.stack([(v if isinstance(v, torch.Tensor) else torch.tensor(float(v))).to(device=device, dtype=torch.float64).reshape(()) for v in step_size])
    return None


def _execute_dev(terms: Sequence[Term], x: torch.Tensor, n_out: int, scale: torch.Tensor) -> torch.Tensor:
    """``_execute`` with the per-axis scales read on the device: term.alpha carries the sign only."""
    spatial = tuple(x.shape[1:])
    out = torch.empty((n_out,) + spatial, dtype=x.dtype, device=x.device)
    if x.numel() == 0:
        return out
    m = _match_fused(terms, x.shape[0], n_out, spatial) if len(spatial) == 3 else None
    if m is not None and all(t.alpha == 1.0 for t in m[1]):
        kind, axis_terms = m
        plans = tuple(t.plan(spatial[k]) for k, t in enumerate(axis_terms))
        xs = x[0] if kind in ("laplacian", "gradient") else [x[0], x[1], x[2]]
        outs = out[0] if kind in ("laplacian", "divergence") else [out[0], out[1], out[2]]
        _cabi.fused3d_dev_alpha(kind, plans, xs, outs, scale.contiguous())
        for co in range(3 if kind in ("gradient", "curl") else 1, n_out):
            out[co].zero_()
        return out
    written = [False] * n_out
    for t in terms:
        _cabi.axis_apply_dev_alpha(t.plan(spatial[t.axis]), x[t.ci], out[t.co], t.axis, scale[t.axis : t.axis + 1], sign=t.alpha, accumulate=written[t.co])
        written[t.co] = True
    for co, w in enumerate(written):
        if not w:
            out[co].zero_()
    return out


class _TracedStencilOp(torch.autograd.Function):
    """A stencil operator whose step sizes are device values: out = sum_terms sign * h_axis**(-d) * D x.  Backward gives
    the cotangent of the array (transposed terms, like ``_StencilOp``) AND of ``step_size`` -- jax.grad through the
    reference yields both: d/dh [h**(-d) D x] = -d / h * (the term itself)."""

    @staticmethod
    def forward(ctx, x, h, terms, n_out):
        d_axis = {t.axis: t.derivative for t in terms}
        hd = h.detach()
        scale = torch.stack([torch.pow(hd[k], -float(d_axis.get(k, 0))) for k in range(hd.numel())])  # 1 / h**d per axis, on the device
        ctx.terms, ctx.n_in = terms, x.shape[0]
        ctx.save_for_backward(x, h.detach(), scale)
        return _execute_dev(terms, x, n_out, scale)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, h, scale = ctx.saved_tensors
        g = g.contiguous()
        gx = gh = None
        if ctx.needs_input_grad[0]:
            gx = _execute_dev(tuple(t.t() for t in ctx.terms), g, ctx.n_in, scale)
        if ctx.needs_input_grad[1]:
            gh = torch.zeros_like(h)
            tmp = torch.empty_like(x[0])
            for t in ctx.terms:  # one pass per term: <g, term> (only when the step sizes want a gradient)
                _cabi.axis_apply_dev_alpha(t.plan(x.shape[1 + t.axis]), x[t.ci], tmp, t.axis, scale[t.axis : t.axis + 1], sign=t.alpha)
                gh[t.axis] += (g[t.co].double() * tmp.double()).sum() * (-float(t.derivative)) / h[t.axis]
        return gx, gh, None, None


def _run_traced(x_fields: torch.Tensor, h: torch.Tensor, terms, n_out: int) -> torch.Tensor:
    terms = tuple(terms)
    for t in terms:
        t.plan(x_fields.shape[1 + t.axis])
    return _TracedStencilOp.apply(x_fields, h, terms, n_out)


def _norm_axis(axis: int, ndim: int) -> int:
    if not -ndim <= axis < ndim:
        raise ValueError(f"axis {axis} is out of bounds for array of dimension {ndim}")
    return axis % ndim


_validated: dict = {}
_plan_mod._dependent_caches.append(_validated.clear)


def _run(x_fields: torch.Tensor, terms, n_out: int) -> torch.Tensor:
    if type(terms) is not tuple:
        terms = tuple(terms)
    vkey = (id(terms), x_fields.shape)
    if _validated.get(vkey) is not terms:
        for t in terms:  # validate sizes/methods up front so errors match the reference's ValueError
            t.plan(x_fields.shape[1 + t.axis])
        if len(_validated) > 256:
            _validated.clear()
        _validated[vkey] = terms  # (keeps the tuple alive, so its id cannot be reused by another term list)
    if not x_fields.requires_grad or not torch.is_grad_enabled():
        return _execute(terms, x_fields, n_out)  # skip the autograd.Function round trip
    return _StencilOp.apply(x_fields, terms, n_out)


_terms_memo: dict = {}
_plan_mod._dependent_caches.append(_terms_memo.clear)


def _memo_terms(op: str, nd: int, accuracy, step_size, method, build):
    """Term lists are pure functions of (operator, rank, accuracy, step_size, method): repeated calls
    with plain Python scalars / tuples (the common case) skip validation and construction."""
    def plain(v):
        return type(v) in (int, float) or (type(v) is tuple and all(type(e) in (int, float) for e in v))

    if not (plain(accuracy) and plain(step_size) and type(method) is str):
        return build()  # arrays / tensors: their values may change, build every time
    key = (op, nd, accuracy, step_size, method)
    hit = _terms_memo.get(key)
    if hit is None:
        hit = tuple(build())
        if len(_terms_memo) > 256:
            _terms_memo.clear()
        _terms_memo[key] = hit
    return hit


# --------------------------------------------------------------------------------- operators
def difference(
    array,
    *,
    axis: int = 0,
    accuracy: int = 1,
    step_size=1.0,
    derivative: int = 1,
    method: MethodKind = "central",
):
    """Compute the finite difference derivative along a given axis with a given accuracy
    using central difference for interior points and forward/backward difference for boundary points

    Args:
        array: input array
        axis: axis along which to compute the gradient. Default is 0
        accuracy: accuracy order of the gradient. Default is 1
        step_size: step size. Default is 1
        derivative: derivative order of the gradient. Default is 1
        method: the method to use (forward, central, backward). Default is central
    Returns:
        Finite difference derivative along the given axis

    Drop-in for the reference's ``difference`` (finite_diff.py:168-297); raises ``ValueError`` when
    the axis is too short for the stencil or the method is unknown (:292-297).
    """
    x, restore = _to_device(array)
    if x.ndim == 0:
        raise ValueError("difference needs an array with at least one axis")
    ax = _norm_axis(axis, x.ndim)
    ht = _traced_steps(step_size, 1, x.device)
    if ht is not None:  # a device value: the kernels read the scale themselves (no host read, no synchronisation)
        one = torch.ones(1, dtype=torch.float64, device=x.device)
        hv = torch.cat([one.expand(ax), ht[0:1], one.expand(x.ndim - ax - 1)])  # (device ops only: capturable)
        return restore(_run_traced(x.unsqueeze(0), hv, [Term(0, 0, ax, int(derivative), int(accuracy), method, 1.0)], 1).squeeze(0))
    term = Term(0, 0, ax, int(derivative), int(accuracy), method, _alpha(step_size, derivative))
    return restore(_run(x.unsqueeze(0), [term], 1).squeeze(0))  # (a view both ways: select would cost a zero-fill in backward)


def gradient(array, *, accuracy=1, step_size=1, method: MethodKind = "central"):
    """Compute the ∇F of input array where F is a scalar function of x and
    returns vectors of the same shape as x stacked along the first axis.

    Drop-in for the reference's ``gradient`` (finite_diff.py:300-351).  Index notation: dF/dxi
    """
    streamed = _host_streamed("gradient", array, accuracy, step_size, method, 3)
    if streamed is not None:
        return streamed
    x, restore = _to_device(array)
    ht = _traced_steps(step_size, x.ndim, x.device)
    if ht is not None:
        acc = _check_and_return(accuracy, x.ndim, "accuracy")
        return restore(_run_traced(x.unsqueeze(0), ht, [Term(k, 0, k, 1, int(a), method, 1.0) for k, a in enumerate(acc)], x.ndim))

    def build():
        acc = _check_and_return(accuracy, x.ndim, "accuracy")
        steps = _check_and_return(step_size, x.ndim, "step_size")
        return [Term(k, 0, k, 1, int(a), method, _alpha(h, 1)) for k, (a, h) in enumerate(zip(acc, steps))]

    return restore(_run(x.unsqueeze(0), _memo_terms("gradient", x.ndim, accuracy, step_size, method, build), x.ndim))


def jacobian(array, *, accuracy=1, step_size=1, method: MethodKind = "central"):
    """Compute the ∂Fi/∂xj of input array where F is a vector function of x (leading axis).

    Drop-in for the reference's ``jacobian`` (finite_diff.py:354-401): output (C, ndim-1, *spatial).
    """
    x, restore = _to_device(array)
    nd = x.ndim - 1
    accuracy = _check_and_return(accuracy, nd, "accuracy")
    ncomp = x.shape[0]
    ht = _traced_steps(step_size, nd, x.device)
    if ht is not None:
        terms = [Term(c * nd + k, c, k, 1, int(a), method, 1.0) for c in range(ncomp) for k, a in enumerate(accuracy)]
        return restore(_run_traced(x, ht, terms, ncomp * nd).reshape((ncomp, nd) + tuple(x.shape[1:])))
    step_size = _check_and_return(step_size, nd, "step_size")
    terms = [
        Term(c * nd + k, c, k, 1, int(a), method, _alpha(h, 1))
        for c in range(ncomp)
        for k, (a, h) in enumerate(zip(accuracy, step_size))
    ]
    out = _run(x, terms, ncomp * nd)
    return restore(out.reshape((ncomp, nd) + tuple(x.shape[1:])))


def divergence(array, *, accuracy=1, step_size=1, keepdims: bool = True, method: MethodKind = "central"):
    """Compute the ∇.F of input array where F is a vector field whose components
    are the first axis of x and returns a scalar field

    Drop-in for the reference's ``divergence`` (finite_diff.py:404-457).  Index notation: dFi/dxi
    """
    streamed = _host_streamed("divergence", array, accuracy, step_size, method, 1)
    if streamed is not None:
        return streamed[None] if keepdims else streamed
    x, restore = _to_device(array)
    nd = x.ndim - 1
    if x.shape[0] < nd:
        raise IndexError(f"index {x.shape[0]} is out of bounds for axis 0 with size {x.shape[0]}")
    ht = _traced_steps(step_size, nd, x.device)
    if ht is not None:
        acc = _check_and_return(accuracy, nd, "accuracy")
        out = _run_traced(x, ht, [Term(0, k, k, 1, int(a), method, 1.0) for k, a in enumerate(acc)], 1)
        return restore(out if keepdims else out.squeeze(0))

    def build():
        acc = _check_and_return(accuracy, nd, "accuracy")
        steps = _check_and_return(step_size, nd, "step_size")
        # zip() in the reference stops at ndim-1 axes: extra leading components are ignored
        return [Term(0, k, k, 1, int(a), method, _alpha(h, 1)) for k, (a, h) in enumerate(zip(acc, steps))]

    terms = _memo_terms("divergence", nd, accuracy, step_size, method, build)
    out = _run(x, terms, 1)
    return restore(out if keepdims else out.squeeze(0))


def hessian(array, *, accuracy=2, step_size=1, method: MethodKind = "central"):
    """Compute hessian of F: R^m -> R.  Index notation: d2F/dxij

    Drop-in for the reference's ``hessian`` (finite_diff.py:460-529), including its indexing
    ``H[i][j] = F[i + j]`` (:499,510,524-529), which for ndim >= 3 makes key 2 both d2/dx1^2 and
    d2/dx0dx2: results are bug-compatible with the reference.  Off-diagonal
    entries are ``difference(difference(x, ax1), ax2)`` so boundary rows compose as they do there.
    """
    x, restore = _to_device(array)
    nd = x.ndim
    accuracy = _check_and_return(accuracy, nd, "accuracy")
    step_size = _check_and_return(step_size, nd, "step_size")
    # F[key] as the reference's dict ends up (:497-521): the diagonal keys 2k first, then the mixed keys a1 + a2 in
    # `combinations` order, later entries replacing earlier ones (for ndim >= 3, key 2 is d2/dx0dx2, not d2/dx1^2)
    producer = {2 * k: (k,) for k in range(nd)}
    for a1, a2 in combinations(range(nd), 2):
        producer[a1 + a2] = (a1, a2)

    def d1(k):
        return Term(0, 0, k, 1, int(accuracy[k]), method, _alpha(step_size[k], 1))

    def d2(k):
        return Term(0, 0, k, 2, int(accuracy[k]), method, _alpha(step_size[k], 2))

    xf = x.unsqueeze(0)
    if x.requires_grad and torch.is_grad_enabled():
        # differentiable composition: one pass per first derivative that a mixed entry needs (the reference recomputes
        # it per pair), one per surviving table entry, ONE stack
        first, table = {}, {}
        for key, who in producer.items():
            if len(who) == 1:
                table[key] = _run(xf, [d2(who[0])], 1)
            else:
                a1, a2 = who
                if a1 not in first:
                    first[a1] = _run(xf, [d1(a1)], 1)
                table[key] = _run(first[a1], [d1(a2)], 1)
        flat = torch.cat([table[i + j] for i in range(nd) for j in range(nd)], dim=0)
        return restore(flat.reshape((nd, nd) + tuple(x.shape)))
    # no gradient wanted: every surviving entry is computed ONCE, straight into its first slot of the (nd, nd, ...)
    # result and, by the same pass, into the other slots with i + j == key (fdx_axis_apply_dup: the stores are repeated;
    # the reference: nd + nd (nd - 1) differences, each a concatenate, then nd + 1 stacks).  3-D: 7 axis passes, no copies.
    for t in [d2(k) for k in range(nd)] + [d1(k) for k in range(nd)]:
        t.plan(x.shape[t.axis])  # the reference's ValueErrors, before anything is launched
    out = torch.empty((nd, nd) + tuple(x.shape), dtype=x.dtype, device=x.device)
    if nd == 3:
        # ONE pass (fdx_hessian3d: 1 read + 9 writes per point); -2 = not covered (reach > 4, unaligned last axis, tiny n)
        p1 = [d1(k).plan(x.shape[k]) for k in range(3)]
        p2 = [d2(0).plan(x.shape[0]), None, d2(2).plan(x.shape[2])]
        a1 = [_alpha(step_size[k], 1) for k in range(3)]
        a2 = [_alpha(step_size[k], 2) for k in range(3)]
        if _cabi.hessian3d(p1, p2, x, out, a1, a2) == 0:
            return restore(out)
    if nd == 2:
        p1 = [d1(k).plan(x.shape[k]) for k in range(2)]
        p2 = [d2(k).plan(x.shape[k]) for k in range(2)]
        if _cabi.hessian2d(p1, p2, x, out, [_alpha(step_size[k], 1) for k in range(2)], [_alpha(step_size[k], 2) for k in range(2)]) == 0:
            return restore(out)
    first = {}
    for key, who in producer.items():
        slots = [(i, key - i) for i in range(nd) if 0 <= key - i < nd]
        dsts = [out[sl] for sl in slots]  # (at most three: H[i][j] with i + j == key)
        if len(who) == 1:
            t = d2(who[0])
            _cabi.axis_apply_dup(t.plan(x.shape[t.axis]), x, dsts, t.axis, t.alpha)
        else:
            a1, a2 = who
            if a1 not in first:
                t = d1(a1)
                first[a1] = torch.empty_like(x)
                _cabi.axis_apply(t.plan(x.shape[a1]), x, first[a1], a1, t.alpha)
            t = d1(a2)
            _cabi.axis_apply_dup(t.plan(x.shape[a2]), first[a1], dsts, a2, t.alpha)
    return restore(out)


def laplacian(array, *, accuracy=1, step_size=1, method: MethodKind = "central"):
    """Compute the ΔF of input array.  Index notation: d2F/dxi2

    Drop-in for the reference's ``laplacian`` (finite_diff.py:532-575).
    """
    streamed = _host_streamed("laplacian", array, accuracy, step_size, method, 1)
    if streamed is not None:
        return streamed
    x, restore = _to_device(array)
    ht = _traced_steps(step_size, x.ndim, x.device)
    if ht is not None:
        acc = _check_and_return(accuracy, x.ndim, "accuracy")
        return restore(_run_traced(x.unsqueeze(0), ht, [Term(0, 0, k, 2, int(a), method, 1.0) for k, a in enumerate(acc)], 1).squeeze(0))

    def build():
        acc = _check_and_return(accuracy, x.ndim, "accuracy")
        steps = _check_and_return(step_size, x.ndim, "step_size")
        return [Term(0, 0, k, 2, int(a), method, _alpha(h, 2)) for k, (a, h) in enumerate(zip(acc, steps))]

    return restore(_run(x.unsqueeze(0), _memo_terms("laplacian", x.ndim, accuracy, step_size, method, build), 1).squeeze(0))


def curl(array, *, accuracy=1, step_size=1, method: MethodKind = "central", keepdims: bool = True):
    """Compute the ∇×F of input array where F is a vector field whose components
    are the first axis of x and returns a vector field.  Index notation: εijk dFk/dxj

    Drop-in for the reference's ``curl`` (finite_diff.py:671-741): 3-D fields need shape (3, ...)
    (``keepdims`` ignored, as there), 2-D fields shape (2, ...); anything else raises ``ValueError``.
    """
    streamed = _host_streamed("curl", array, accuracy, step_size, method, 3)
    if streamed is not None:
        return streamed
    x, restore = _to_device(array)
    nd = x.ndim - 1
    accuracy = _check_and_return(accuracy, nd, "accuracy")
    ht = _traced_steps(step_size, nd, x.device) if nd >= 1 else None
    if ht is None:
        step_size = _check_and_return(step_size, nd, "step_size")
        run = _run
    else:
        run = lambda xf, terms, n_out: _run_traced(xf, ht, terms, n_out)  # noqa: E731

    def term(co, ci, ax, sign):
        return Term(co, ci, ax, 1, int(accuracy[ax]), method, sign * (1.0 if ht is not None else _alpha(step_size[ax], 1)))

    if x.ndim == 4 and x.shape[0] == 3:
        terms = [
            term(0, 2, 1, +1), term(0, 1, 2, -1),
            term(1, 0, 2, +1), term(1, 2, 0, -1),
            term(2, 1, 0, +1), term(2, 0, 1, -1),
        ]  # fmt: skip
        return restore(run(x, terms, 3))
    if x.ndim == 3 and x.shape[0] == 2:
        out = run(x, [term(0, 1, 0, +1), term(0, 0, 1, -1)], 1)
        return restore(out if keepdims else out.squeeze(0))
    raise ValueError(
        f"`curl` is only implemented for 2D and 3D vector fields, got array.ndim={x.ndim}"
        "for 2D vector fields, the leading axis must have a shape=2, "
        "for 3D vector fields, the leading axis must have a shape=3"
    )
