"""jax.ffi registration of the fdx_b200 kernels (the binding the north star names).

NOT exercised in this repository's image: JAX is not installed there -- this module and
``fdx_jax_ffi.cc`` have never been imported / compiled.  Importing it without JAX raises ImportError; nothing else in
the package depends on it.  What it binds is tested: the C entry points below are the ones ``tests/`` drive through
ctypes, the ``*_dev_alpha`` ones under CUDA-graph capture.

With JAX present it (1) builds ``fdx_jax_ffi.cc`` against ``jax.ffi.include_dir()``, (2) registers the handlers for
platform "CUDA" (execute stage + a shared instantiate stage that creates the device plans), and (3) exposes
``difference``, ``gradient``, ``divergence``, ``laplacian`` and ``curl`` with the reference's signatures, each wrapped
in ``jax.custom_vjp`` whose backward rule applies the transposed plans (boundary rows included) and returns the
cotangent of ``step_size`` as well.  3-D grids use the fused handlers; everything else is composed from ``difference``
exactly as the reference composes it (``jacobian`` and ``hessian`` are such compositions there, finite_diff.py:392-401,
492-529, and stay compositions here).

``step_size``:
* a Python / NumPy number is static -> ``fdx_*3d_*_static_ffi``: the single-pass fused kernels;
* a traced value -> ``alpha`` stays a device operand and the ``*_dev_alpha`` entry points read it on the device
  (no host read, no stream synchronisation; the 3-D operators then run as axis passes, see include/fdx_b200.h).

Inputs that are not float32 / float64 are cast to float32 first (the reference's promotion with x64 off).
"""
from __future__ import annotations

import ctypes
import functools
import os
import subprocess

import numpy as np

import jax  # noqa: F401  (ImportError here is the documented behaviour without JAX)
import jax.numpy as jnp

from .. import _build
from ..plan import AxisPlan, axis_plan, axis_plan_t
from ..utils import _check_and_return

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(os.path.dirname(_HERE), "_lib", "libfdx_jax_ffi.so")
_KINDS = ("laplacian", "divergence", "gradient", "curl")


def build() -> str:
    lib = _build.build()
    if not os.path.exists(_SO):
        subprocess.run(
            ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", f"-I{jax.ffi.include_dir()}", "-I/usr/local/cuda/include",
             f"-I{os.path.join(os.path.dirname(os.path.dirname(_HERE)), 'include')}", os.path.join(_HERE, "fdx_jax_ffi.cc"),
             f"-L{os.path.dirname(lib)}", "-lfdx_b200", "-lcudart", f"-Wl,-rpath,{os.path.dirname(lib)}", "-o", _SO],
            check=True,
        )
    return _SO


@functools.lru_cache(maxsize=None)
def _register():
    so = ctypes.CDLL(build())
    inst = jax.ffi.pycapsule(so.fdx_instantiate_ffi)
    names = ["fdx_axis_apply_f32_ffi", "fdx_axis_apply_f64_ffi"]
    for k in _KINDS:
        for sfx in ("f32", "f64"):
            names += [f"fdx_{k}3d_{sfx}_ffi", f"fdx_{k}3d_{sfx}_static_ffi"]
    for name in names:
        jax.ffi.register_ffi_target(name, {"instantiate": inst, "execute": jax.ffi.pycapsule(getattr(so, name))}, platform="CUDA")


def _plan_attrs(plans) -> dict:
    """Flattened plan attributes (see KeysFrom in fdx_jax_ffi.cc)."""
    plans = list(plans)
    return dict(
        plan_shape=np.array([v for p in plans for v in (p.n, p.lo, p.hi, p.nl, p.wl, p.nr, p.wr)], dtype=np.int32),
        main_taps=np.concatenate([np.asarray(p.main, dtype=np.float64) for p in plans]),
        band_l=np.concatenate([np.asarray(p.band_l, dtype=np.float64).reshape(-1) for p in plans]),
        band_r=np.concatenate([np.asarray(p.band_r, dtype=np.float64).reshape(-1) for p in plans]),
    )


def _floating(x):
    x = jnp.asarray(x)
    return x if x.dtype in (jnp.float32, jnp.float64) else x.astype(jnp.float32)


def _sfx(x) -> str:
    return "f64" if x.dtype == jnp.float64 else "f32"


def _is_static(step_size) -> bool:
    vals = step_size if isinstance(step_size, tuple) else (step_size,)
    return all(isinstance(v, (int, float, np.integer, np.floating)) for v in vals)


# ------------------------------------------------------------------------------------------------ difference
def _axis_call(x, alpha, plan: AxisPlan, axis: int):
    _register()
    call = jax.ffi.ffi_call(f"fdx_axis_apply_{_sfx(x)}_ffi", jax.ShapeDtypeStruct(x.shape, x.dtype), vmap_method="sequential")
    return call(x, jnp.asarray(alpha, dtype=jnp.float64).reshape(1), axis=np.int64(axis), **_plan_attrs([plan]))


def _difference(x, alpha, n, axis, derivative, accuracy, method):
    fwd = axis_plan(n, derivative, accuracy, method, axis)
    bwd = axis_plan_t(n, derivative, accuracy, method)

    @jax.custom_vjp
    def op(x, alpha):
        return _axis_call(x, alpha, fwd, axis)

    def op_fwd(x, alpha):
        out = _axis_call(x, alpha, fwd, axis)
        return out, (alpha, out)

    def op_bwd(res, g):
        alpha, out = res
        # transposed stencil, boundary rows included; the operator is linear in alpha: d/d(alpha) = <g, out> / alpha
        return _axis_call(g, alpha, bwd, axis), jnp.vdot(g, out) / alpha

    op.defvjp(op_fwd, op_bwd)
    return op(x, alpha)


@functools.partial(jax.jit, static_argnames=("accuracy", "axis", "derivative", "method"))
def difference(array, *, axis=0, accuracy=1, step_size=1.0, derivative=1, method="central"):
    """Drop-in for ``finitediffx.difference`` (finite_diff.py:168-297) on the CUDA platform."""
    x = _floating(array)
    axis = axis % x.ndim
    return _difference(x, 1.0 / (step_size**derivative), x.shape[axis], axis, derivative, accuracy, method)


# ------------------------------------------------------------------------------------------------ fused 3-D operators
_T = {"laplacian": "laplacian", "gradient": "divergence", "divergence": "gradient", "curl": "curl"}  # kind of the adjoint


def _fused(kind: str, x, alpha, static_alpha, fwd_plans, bwd_plans, out_shape):
    """One fused 3-D operator with its custom_vjp.  alpha: f64[3] device operand (traced step sizes) or None together
    with static_alpha (three Python floats)."""
    _register()

    def call(k, v, a, plans, shape, sign=1.0):
        if static_alpha is not None:
            c = jax.ffi.ffi_call(f"fdx_{k}3d_{_sfx(v)}_static_ffi", jax.ShapeDtypeStruct(shape, v.dtype), vmap_method="sequential")
            return c(v, alpha=np.asarray(static_alpha, dtype=np.float64) * sign, **_plan_attrs(plans))
        c = jax.ffi.ffi_call(f"fdx_{k}3d_{_sfx(v)}_ffi", jax.ShapeDtypeStruct(shape, v.dtype), vmap_method="sequential")
        return c(v, a * sign, **_plan_attrs(plans))

    @jax.custom_vjp
    def op(x, a):
        return call(kind, x, a, fwd_plans, out_shape)

    def op_fwd(x, a):
        return call(kind, x, a, fwd_plans, out_shape), (x, a)

    def op_bwd(res, g):
        x, a = res
        # the adjoint of curl is curl with transposed plans and negated scales; of gradient, divergence (and back)
        gx = call(_T[kind], g, a, bwd_plans, x.shape, sign=-1.0 if kind == "curl" else 1.0)
        if static_alpha is not None:
            return gx, None
        # d out / d alpha_k = (the axis-k part of the operator) / alpha_k: one axis pass per scale, only on this branch
        ga = []
        for k in range(3):
            ek = jnp.zeros(3, dtype=jnp.float64).at[k].set(1.0)
            ga.append(jnp.vdot(g, call(kind, x, ek, fwd_plans, out_shape)))
        return gx, jnp.stack(ga)

    op.defvjp(op_fwd, op_bwd)
    return op(x, alpha)


def _three(value, name):
    return _check_and_return(value, 3, name)


def _operator(kind: str, array, accuracy, step_size, method, derivative: int):
    x = _floating(array)
    spatial = x.shape[-3:]
    acc = _three(accuracy, "accuracy")
    static = _is_static(step_size)
    steps = _three(step_size, "step_size") if static else jnp.broadcast_to(jnp.asarray(step_size, dtype=jnp.float64).reshape(-1), (3,))
    fwd = [axis_plan(spatial[k], derivative, int(acc[k]), method, k) for k in range(3)]
    bwd = [axis_plan_t(spatial[k], derivative, int(acc[k]), method) for k in range(3)]
    out_shape = {"laplacian": spatial, "divergence": spatial, "gradient": (3,) + spatial, "curl": (3,) + spatial}[kind]
    if static:
        return _fused(kind, x, None, tuple(1.0 / float(h) ** derivative for h in steps), fwd, bwd, out_shape)
    return _fused(kind, x, 1.0 / steps**derivative, None, fwd, bwd, out_shape)


@functools.partial(jax.jit, static_argnames=("accuracy", "method"))
def laplacian(array, *, accuracy=1, step_size=1, method="central"):
    """Drop-in for ``finitediffx.laplacian`` (finite_diff.py:532-575); 3-D grids run fused, others are composed."""
    x = _floating(array)
    if x.ndim == 3:
        return _operator("laplacian", x, accuracy, step_size, method, 2)
    acc, steps = _check_and_return(accuracy, x.ndim, "accuracy"), _check_and_return(step_size, x.ndim, "step_size")
    return sum(difference(x, axis=k, accuracy=acc[k], step_size=steps[k], derivative=2, method=method) for k in range(x.ndim))


@functools.partial(jax.jit, static_argnames=("accuracy", "method"))
def gradient(array, *, accuracy=1, step_size=1, method="central"):
    """Drop-in for ``finitediffx.gradient`` (finite_diff.py:300-351)."""
    x = _floating(array)
    if x.ndim == 3:
        return _operator("gradient", x, accuracy, step_size, method, 1)
    acc, steps = _check_and_return(accuracy, x.ndim, "accuracy"), _check_and_return(step_size, x.ndim, "step_size")
    return jnp.stack([difference(x, axis=k, accuracy=acc[k], step_size=steps[k], method=method) for k in range(x.ndim)])


@functools.partial(jax.jit, static_argnames=("accuracy", "method", "keepdims"))
def divergence(array, *, accuracy=1, step_size=1, keepdims=True, method="central"):
    """Drop-in for ``finitediffx.divergence`` (finite_diff.py:404-457)."""
    x = _floating(array)
    if x.ndim == 4 and x.shape[0] >= 3:
        out = _operator("divergence", x[:3], accuracy, step_size, method, 1)
    else:
        nd = x.ndim - 1
        acc, steps = _check_and_return(accuracy, nd, "accuracy"), _check_and_return(step_size, nd, "step_size")
        out = sum(difference(x[k], axis=k, accuracy=acc[k], step_size=steps[k], method=method) for k in range(nd))
    return out[None] if keepdims else out


@functools.partial(jax.jit, static_argnames=("accuracy", "method", "keepdims"))
def curl(array, *, accuracy=1, step_size=1, method="central", keepdims=True):
    """Drop-in for ``finitediffx.curl`` (finite_diff.py:671-741)."""
    x = _floating(array)
    if x.ndim == 4 and x.shape[0] == 3:
        return _operator("curl", x, accuracy, step_size, method, 1)
    if x.ndim == 3 and x.shape[0] == 2:
        acc, steps = _check_and_return(accuracy, 2, "accuracy"), _check_and_return(step_size, 2, "step_size")
        d = lambda c, ax: difference(x[c], axis=ax, accuracy=acc[ax], step_size=steps[ax], method=method)  # noqa: E731
        out = d(1, 0) - d(0, 1)
        return out[None] if keepdims else out
    raise ValueError(
        f"`curl` is only implemented for 2D and 3D vector fields, got array.ndim={x.ndim}"
        "for 2D vector fields, the leading axis must have a shape=2, "
        "for 3D vector fields, the leading axis must have a shape=3"
    )
