"""jax.ffi registration of the fdx_b200 kernels (the binding the north star names).

NOT exercised in this repository's image: JAX is not installed there.  Importing
this module without JAX raises ImportError; nothing else in the package depends on it.  With JAX
present it (1) builds ``fdx_jax_ffi.cc`` against ``jax.ffi.include_dir()``, (2) registers the
handlers for platform "CUDA", and (3) exposes ``difference`` with the reference's signature wrapped
in ``jax.custom_vjp`` whose backward rule applies the transposed plan -- the same plans
(`finitediffx_b200.plan`) and the same C ABI the torch/ctypes binding uses.
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

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(os.path.dirname(_HERE), "_lib", "libfdx_jax_ffi.so")


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
    for name in ("fdx_axis_apply_f32_ffi", "fdx_axis_apply_f64_ffi"):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(so, name)), platform="CUDA")


def _plan_attrs(p: AxisPlan) -> dict:
    return dict(
        plan_shape=np.array([p.n, p.lo, p.hi, p.nl, p.wl, p.nr, p.wr], dtype=np.int32),
        main_taps=np.asarray(p.main, dtype=np.float64),
        band_l=np.asarray(p.band_l, dtype=np.float64).reshape(-1),
        band_r=np.asarray(p.band_r, dtype=np.float64).reshape(-1),
    )


def _apply(x, alpha, plan: AxisPlan, axis: int):
    _register()
    name = "fdx_axis_apply_f64_ffi" if x.dtype == jnp.float64 else "fdx_axis_apply_f32_ffi"
    call = jax.ffi.ffi_call(name, jax.ShapeDtypeStruct(x.shape, x.dtype), vmap_method="sequential")
    return call(x, jnp.asarray(alpha, dtype=jnp.float64).reshape(1), axis=np.int64(axis), **_plan_attrs(plan))


@functools.partial(jax.jit, static_argnames=("accuracy", "axis", "derivative", "method"))
def difference(array, *, axis=0, accuracy=1, step_size=1.0, derivative=1, method="central"):
    """Drop-in for ``finitediffx.difference`` (finite_diff.py:168-297) on the CUDA platform."""
    axis = axis % array.ndim
    n = array.shape[axis]
    fwd = axis_plan(n, derivative, accuracy, method, axis)
    bwd = axis_plan_t(n, derivative, accuracy, method)

    @jax.custom_vjp
    def op(x, alpha):
        return _apply(x, alpha, fwd, axis)

    def op_fwd(x, alpha):
        out = _apply(x, alpha, fwd, axis)
        return out, (alpha, out)

    def op_bwd(res, g):
        alpha, out = res
        # transposed stencil, boundary rows included; d/d(alpha) = <g, out> / alpha
        return _apply(g, alpha, bwd, axis), jnp.vdot(g, out) / alpha

    op.defvjp(op_fwd, op_bwd)
    x = array if jnp.issubdtype(array.dtype, jnp.floating) else array.astype(jnp.float32)
    return op(x, 1.0 / (step_size**derivative))
