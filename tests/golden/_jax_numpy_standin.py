"""A NumPy stand-in for the handful of ``jax`` entry points the reference's array path calls.

TEST INFRASTRUCTURE.  ``jax``/``jaxlib`` are not installed in the build image, so
the reference cannot be imported as is.  ``make_golden.py`` installs this module as ``jax`` /
``jax.numpy`` / ``jax.tree_util`` / ``jax.lax`` in ``sys.modules`` and then imports the
reference's *unmodified* source from /root/reference, so that the golden vectors under
``tests/golden/`` are produced by the reference's own Python control flow (slice bounds, row
regions, summation order, dict keying) rather than by our restatement of it.

What is emulated, and how faithfully:
  * dtype widths follow ``X64`` (the ``jax_enable_x64`` flag): Python ints/floats and
    ``arange`` become int32/float32 when it is off, int64/float64 when on;
  * JAX's promotion lattice for the cases that occur (int (+) float -> that float width, no
    silent widening to float64 by integer operands; Python scalars are weak);
  * ``jax.jit`` turns non-static Python scalars into 0-d arrays (tracers are ``jax.Array``s,
    which is what routes a float ``step_size`` through ``jnp.repeat`` in utils.py:28-29);
  * ``jax.lax.slice_in_dim`` raises on out-of-range slices like XLA does.
Not emulated: XLA's own rounding inside ``jnp.linalg.inv`` (LAPACK getrf/getri via NumPy is
used instead) and XLA's fusion/FMA contraction choices on CPU.
"""

from __future__ import annotations

import functools
import inspect
import sys
import types

import numpy as np

X64 = True  # toggled by make_golden.enable_x64(...)


def _ft():
    return np.float64 if X64 else np.float32


def _it():
    return np.int64 if X64 else np.int32


class Array(np.ndarray):
    """ndarray subclass applying JAX-style promotion inside ufuncs (incl. matmul)."""

    def __new__(cls, value, dtype=None):
        return np.asarray(value, dtype=dtype).view(cls)

    def __array_finalize__(self, obj):
        pass

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if ufunc is np.power and method == "__call__" and isinstance(inputs[1], int) and inputs[1] >= 0 and out is None:
            # jnp power with a static Python int exponent lowers to lax.integer_pow
            base, y, acc = inputs[0], inputs[1], None
            while y > 0:
                if y & 1:
                    acc = base if acc is None else acc * base
                y >>= 1
                if y > 0:
                    base = base * base
            return _wrap(np.ones_like(inputs[0])) if acc is None else acc
        raw, float_widths, has_int_only = [], [], True
        for v in inputs:
            if isinstance(v, np.ndarray):
                a = v.view(np.ndarray)
                if np.issubdtype(a.dtype, np.floating):
                    float_widths.append(a.dtype)
                    has_int_only = False
                raw.append(a)
            elif isinstance(v, np.generic):
                a = np.asarray(v)
                if np.issubdtype(a.dtype, np.floating):
                    float_widths.append(a.dtype)
                    has_int_only = False
                raw.append(a)
            else:
                if isinstance(v, float):
                    has_int_only = False
                raw.append(v)  # weak Python scalar
        if float_widths:
            target = max(float_widths, key=lambda d: d.itemsize)
            raw = [a.astype(target) if isinstance(a, np.ndarray) and a.dtype != np.bool_ else a for a in raw]
        elif not has_int_only:
            # int arrays with a weak Python float -> default float width
            raw = [a.astype(_ft()) if isinstance(a, np.ndarray) and a.dtype != np.bool_ else a for a in raw]
        elif ufunc in (np.true_divide,):
            raw = [a.astype(_ft()) if isinstance(a, np.ndarray) else a for a in raw]
        if out is not None:
            kwargs["out"] = tuple(o.view(np.ndarray) if isinstance(o, np.ndarray) else o for o in out)
        with np.errstate(over="ignore"):
            res = getattr(ufunc, method)(*raw, **kwargs)
        if isinstance(res, tuple):
            return tuple(_wrap(r) for r in res)
        return _wrap(res)

    def __iter__(self):
        # iterating a jax array yields 0-d jax arrays (strongly typed), not Python scalars
        for i in range(self.shape[0]):
            yield self[i]

    def __getitem__(self, item):
        return _wrap(np.ndarray.__getitem__(self.view(np.ndarray), item))

    def flatten(self):
        return _wrap(self.view(np.ndarray).flatten())

    def __hash__(self):  # tracers are not hashable in jax either, but dict keys never see them
        return id(self)


def _wrap(value):
    if isinstance(value, np.ndarray):
        return value.view(Array)
    if isinstance(value, np.generic):
        return np.asarray(value).view(Array)
    return value


def _from_python(value):
    """What tracing does to a non-static argument leaf."""
    if isinstance(value, bool):
        return value
    if isinstance(value, int):
        return Array(value, dtype=_it())
    if isinstance(value, float):
        return Array(value, dtype=_ft())
    if isinstance(value, (np.ndarray, np.generic)):
        a = np.asarray(value)
        if not X64:
            if a.dtype == np.float64:
                a = a.astype(np.float32)
            elif a.dtype == np.int64:
                a = a.astype(np.int32)
        return a.view(Array)
    if isinstance(value, tuple):
        return tuple(_from_python(v) for v in value)
    if isinstance(value, list):
        return [_from_python(v) for v in value]
    return value


def jit(fun=None, *, static_argnums=(), static_argnames=()):
    if fun is None:
        return functools.partial(jit, static_argnums=static_argnums, static_argnames=static_argnames)
    if isinstance(static_argnums, int):
        static_argnums = (static_argnums,)
    if isinstance(static_argnames, str):
        static_argnames = (static_argnames,)
    sig = inspect.signature(fun)
    names = list(sig.parameters)

    @functools.wraps(fun)
    def wrapped(*args, **kwargs):
        new_args = [a if i in static_argnums or (i < len(names) and names[i] in static_argnames) else _from_python(a) for i, a in enumerate(args)]
        new_kwargs = {k: (v if k in static_argnames else _from_python(v)) for k, v in kwargs.items()}
        return fun(*new_args, **new_kwargs)

    return wrapped


# ------------------------------------------------------------------------- jax.numpy
def _array(value, dtype=None):
    if isinstance(value, (tuple, list)) and all(isinstance(v, np.ndarray) and v.ndim == 0 for v in value) and value:
        a = np.stack([np.asarray(v) for v in value])
    else:
        a = np.asarray(value)
    if dtype is None:
        if a.dtype == np.float64 and not X64:
            a = a.astype(np.float32)
        if a.dtype == np.int64 and not X64:
            a = a.astype(np.int32)
    else:
        a = a.astype(dtype)
    return a.view(Array)


def _arange(*args, **kwargs):
    a = np.arange(*args, **kwargs)
    if np.issubdtype(a.dtype, np.integer):
        a = a.astype(_it())
    elif np.issubdtype(a.dtype, np.floating):
        a = a.astype(_ft())
    return a.view(Array)


def _where(cond, x, y):
    xa, ya = np.asarray(x), np.asarray(y)
    if isinstance(x, int) and isinstance(y, int):
        return np.where(np.asarray(cond), xa.astype(_it()), ya.astype(_it())).view(Array)
    return _wrap(np.where(np.asarray(cond), xa, ya))


def _inv(a):
    a = np.asarray(a)
    if not np.issubdtype(a.dtype, np.floating):
        a = a.astype(_ft())
    return np.linalg.inv(a).astype(a.dtype).view(Array)


def _slice_in_dim(operand, start_index, limit_index, stride=1, axis=0):
    n = operand.shape[axis]
    start_index, limit_index = int(start_index), int(limit_index)
    if not (0 <= start_index <= limit_index <= n):
        raise TypeError(
            f"slice start_indices/limit_indices ({start_index}, {limit_index}) out of range for dimension of size {n}"
        )
    sl = [slice(None)] * operand.ndim
    sl[axis] = slice(start_index, limit_index, stride)
    return _wrap(np.asarray(operand)[tuple(sl)])


def _concatenate(arrays, axis=0):
    arrays = [np.asarray(a) for a in arrays]
    dt = np.result_type(*[a.dtype for a in arrays])
    return np.concatenate([a.astype(dt) for a in arrays], axis=axis).view(Array)


def _stack(arrays, axis=0):
    arrays = [np.asarray(a) for a in arrays]
    return np.stack(arrays, axis=axis).view(Array)


def install():
    """Register the stand-in under the ``jax`` names; returns the fake top-level module."""
    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    lax = types.ModuleType("jax.lax")
    jtu = types.ModuleType("jax.tree_util")
    linalg = types.ModuleType("jax.numpy.linalg")
    errors = types.ModuleType("jax.errors")

    jax.Array = Array
    jax.jit = jit
    jax.lax, jax.numpy, jax.tree_util, jax.errors = lax, jnp, jtu, errors
    jax.custom_jvp = lambda f, *a, **k: f
    jax.vmap = lambda f, *a, **k: f
    jax.__version__ = "0.0-numpy-standin"
    lax.slice_in_dim = _slice_in_dim
    jtu.PyTreeDef = object
    errors.TracerArrayConversionError = TypeError

    jnp.ndarray = Array
    jnp.array = _array
    jnp.asarray = _array
    jnp.arange = _arange
    jnp.where = _where
    jnp.repeat = lambda a, repeats, axis=None: _wrap(np.repeat(np.asarray(a), int(repeats), axis=axis))
    jnp.concatenate = _concatenate
    jnp.stack = _stack
    jnp.expand_dims = lambda a, axis: _wrap(np.expand_dims(np.asarray(a), axis))
    jnp.linalg = linalg
    linalg.inv = _inv

    for name, mod in {
        "jax": jax,
        "jax.numpy": jnp,
        "jax.lax": lax,
        "jax.tree_util": jtu,
        "jax.numpy.linalg": linalg,
        "jax.errors": errors,
    }.items():
        sys.modules[name] = mod
    return jax


def set_x64(flag: bool):
    global X64
    X64 = bool(flag)
