"""ctypes binding of libfdx_b200.so (the C ABI declared in include/fdx_b200.h).

This is the testable stand-in for the ``jax.ffi`` registration the north star names (JAX is not
installed in this image; INTEGRATION.md shows the XLA-FFI shim over the same symbols).  It only
marshals raw device pointers, sizes and the current CUDA stream: torch is used for device memory
and streams, nothing else.

There is no CPU fallback: if the shared library is missing, ``load()`` raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Sequence

import numpy as np
import torch

from .plan import MAX_TAPS, AxisPlan

# FDX_B200_LIB: another build of the same library (A/B timing of kernel variants, tools/ab_time.sh)
_LIB_PATH = os.environ.get("FDX_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libfdx_b200.so")
_lib = None
_lock = threading.Lock()

MAX_BAND = 64  # FDX_MAX_BAND


class FdxError(RuntimeError):
    pass


class _AxisDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("lo", C.c_int32),
        ("hi", C.c_int32),
        ("main_taps", C.c_double * MAX_TAPS),
        ("nl", C.c_int32),
        ("wl", C.c_int32),
        ("nr", C.c_int32),
        ("wr", C.c_int32),
        ("band_l", C.POINTER(C.c_double)),
        ("band_r", C.POINTER(C.c_double)),
    ]


class Slab(C.Structure):
    _fields_ = [
        ("z_begin", C.c_int32),
        ("z_end", C.c_int32),
        ("in_first", C.c_int32),
        ("in_planes", C.c_int32),
        ("out_first", C.c_int32),
        ("z_begin2", C.c_int32),  # optional second range of output planes (the two shells of a slab in one launch)
        ("z_end2", C.c_int32),
    ]


EXPORTS = (
    "fdx_plan_create",
    "fdx_plan_destroy",
    "fdx_axis_apply_f32",
    "fdx_axis_apply_f64",
    "fdx_laplacian3d_f32",
    "fdx_laplacian3d_f64",
    "fdx_divergence3d_f32",
    "fdx_divergence3d_f64",
    "fdx_gradient3d_f32",
    "fdx_gradient3d_f64",
    "fdx_curl3d_f32",
    "fdx_curl3d_f64",
    "fdx_version",
    "fdx_error_string",
    "fdx_launch_count",
    "fdx_last_kernel",
    "fdx_reload_env",
    "fdx_axis_apply_dev_alpha_f32",
    "fdx_axis_apply_dev_alpha_f64",
    "fdx_laplacian3d_dev_alpha_f32",
    "fdx_laplacian3d_dev_alpha_f64",
    "fdx_divergence3d_dev_alpha_f32",
    "fdx_divergence3d_dev_alpha_f64",
    "fdx_gradient3d_dev_alpha_f32",
    "fdx_gradient3d_dev_alpha_f64",
    "fdx_curl3d_dev_alpha_f32",
    "fdx_curl3d_dev_alpha_f64",
    "fdx_laplacian2d_f32",
    "fdx_laplacian2d_f64",
    "fdx_divergence2d_f32",
    "fdx_divergence2d_f64",
    "fdx_gradient2d_f32",
    "fdx_gradient2d_f64",
    "fdx_axis_apply_dup_f32",
    "fdx_axis_apply_dup_f64",
    "fdx_hessian3d_f32",
    "fdx_hessian3d_f64",
    "fdx_hessian2d_f32",
    "fdx_hessian2d_f64",
)


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load libfdx_b200.so; raise loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise FdxError(
                f"{_LIB_PATH} is missing: the CUDA library has not been built. Run "
                "`python -m finitediffx_b200._build` (needs nvcc). There is no CPU fallback."
            )
        lib = C.CDLL(_LIB_PATH)
        if os.environ.get("FDX_B200_LIB"):  # A/B timing against an older build: newer entry points may be missing
            def _missing(name):
                def stub(*a, **k):
                    raise FdxError(f"{name} is not exported by {_LIB_PATH} (an older build selected with FDX_B200_LIB)")

                return stub

            for name in EXPORTS:
                if not hasattr(lib, name):
                    setattr(lib, name, _missing(name))
        vp, i64, dbl, i32 = C.c_void_p, C.c_int64, C.c_double, C.c_int
        lib.fdx_plan_create.argtypes = [C.POINTER(vp), C.POINTER(_AxisDesc)]
        lib.fdx_plan_destroy.argtypes = [vp]
        for sfx in ("f32", "f64"):
            getattr(lib, f"fdx_axis_apply_{sfx}").argtypes = [vp, vp, vp, i64, i64, dbl, i32, vp]
            getattr(lib, f"fdx_axis_apply_dup_{sfx}").argtypes = [vp, vp, vp, vp, vp, i64, i64, dbl, vp]
            getattr(lib, f"fdx_hessian3d_{sfx}").argtypes = [C.POINTER(vp), C.POINTER(vp), vp, vp, C.POINTER(dbl), C.POINTER(dbl), vp]
            getattr(lib, f"fdx_hessian2d_{sfx}").argtypes = [C.POINTER(vp), C.POINTER(vp), vp, vp, C.POINTER(dbl), C.POINTER(dbl), vp]
            getattr(lib, f"fdx_laplacian2d_{sfx}").argtypes = [C.POINTER(vp), vp, vp, C.POINTER(dbl), vp]
            getattr(lib, f"fdx_divergence2d_{sfx}").argtypes = [C.POINTER(vp), C.POINTER(vp), vp, C.POINTER(dbl), vp]
            getattr(lib, f"fdx_gradient2d_{sfx}").argtypes = [C.POINTER(vp), vp, C.POINTER(vp), C.POINTER(dbl), vp]
            getattr(lib, f"fdx_axis_apply_dev_alpha_{sfx}").argtypes = [vp, vp, vp, i64, i64, vp, dbl, i32, vp]
            getattr(lib, f"fdx_laplacian3d_dev_alpha_{sfx}").argtypes = [C.POINTER(vp), vp, vp, vp, vp]
            getattr(lib, f"fdx_divergence3d_dev_alpha_{sfx}").argtypes = [C.POINTER(vp), C.POINTER(vp), vp, vp, vp]
            getattr(lib, f"fdx_gradient3d_dev_alpha_{sfx}").argtypes = [C.POINTER(vp), vp, C.POINTER(vp), vp, vp]
            getattr(lib, f"fdx_curl3d_dev_alpha_{sfx}").argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp, vp]
            getattr(lib, f"fdx_laplacian3d_{sfx}").argtypes = [C.POINTER(vp), vp, vp, C.POINTER(dbl), C.POINTER(Slab), vp]
            getattr(lib, f"fdx_divergence3d_{sfx}").argtypes = [C.POINTER(vp), C.POINTER(vp), vp, C.POINTER(dbl), C.POINTER(Slab), vp]
            getattr(lib, f"fdx_gradient3d_{sfx}").argtypes = [C.POINTER(vp), vp, C.POINTER(vp), C.POINTER(dbl), C.POINTER(Slab), vp]
            getattr(lib, f"fdx_curl3d_{sfx}").argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(dbl), C.POINTER(Slab), vp]
        lib.fdx_error_string.restype = C.c_char_p
        lib.fdx_error_string.argtypes = [i32]
        lib.fdx_launch_count.restype = i64
        lib.fdx_last_kernel.restype = C.c_char_p
        lib.fdx_reload_env.restype = None
        for name in EXPORTS:
            getattr(lib, name)  # AttributeError here means header and library disagree
        _lib = lib
    return _lib


def check(code: int, what: str = "fdx call"):
    if code != 0:
        msg = load().fdx_error_string(code).decode()
        raise FdxError(f"{what} failed with code {code}: {msg}")


def launch_count() -> int:
    return int(load().fdx_launch_count())


def reload_env() -> None:
    """Make the library re-read the FDX_* switches (it snapshots them at first use)."""
    load().fdx_reload_env()


def last_kernel() -> str:
    return load().fdx_last_kernel().decode()


# ------------------------------------------------------------------------------ device plans
class DevicePlan:
    """Device-resident plan handle (fdx_plan_t) built from a host AxisPlan."""

    def __init__(self, plan: AxisPlan, device: torch.device):
        if plan.taps > MAX_TAPS:
            raise FdxError(f"{plan.taps} taps exceed FDX_MAX_TAPS={MAX_TAPS}")
        self.host = plan
        self.device = device
        desc = _AxisDesc()
        desc.n, desc.lo, desc.hi = plan.n, plan.lo, plan.hi
        for k in range(MAX_TAPS):
            desc.main_taps[k] = float(plan.main[k]) if k < plan.taps else 0.0
        desc.nl, desc.wl, desc.nr, desc.wr = plan.nl, plan.wl, plan.nr, plan.wr
        self._bl = np.ascontiguousarray(plan.band_l, dtype=np.float64)
        self._br = np.ascontiguousarray(plan.band_r, dtype=np.float64)
        desc.band_l = self._bl.ctypes.data_as(C.POINTER(C.c_double)) if self._bl.size else None
        desc.band_r = self._br.ctypes.data_as(C.POINTER(C.c_double)) if self._br.size else None
        handle = C.c_void_p()
        with torch.cuda.device(device):
            check(load().fdx_plan_create(C.byref(handle), C.byref(desc)), "fdx_plan_create")
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None) and _lib is not None:
                _lib.fdx_plan_destroy(self.handle)
        except Exception:
            pass


_plan_cache: dict = {}


def device_plan(plan: AxisPlan, device: torch.device) -> DevicePlan:
    """The coefficient/plan cache: one device plan per (host plan, device)."""
    key = (id(plan), device.index if device.index is not None else torch.cuda.current_device())
    hit = _plan_cache.get(key)
    if hit is None or hit.host is not plan:
        hit = DevicePlan(plan, device)
        _plan_cache[key] = hit
    return hit


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _on_device:
    """`with torch.cuda.device(d)` only when d is not already current (the common case is free)."""

    __slots__ = ("ctx",)

    def __init__(self, device):
        idx = device.index
        self.ctx = None if idx is None or idx == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _sfx(t: torch.Tensor) -> str:
    if t.dtype == torch.float32:
        return "f32"
    if t.dtype == torch.float64:
        return "f64"
    raise FdxError(f"unsupported dtype {t.dtype}")


def _require(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise FdxError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise FdxError(f"{name} must be contiguous")


def axis_apply(plan: AxisPlan, x: torch.Tensor, out: torch.Tensor, axis: int, alpha: float, accumulate: bool = False):
    """out (+)= alpha * D_axis x for contiguous CUDA tensors of identical shape."""
    _require(x, "x")
    _require(out, "out")
    assert x.shape == out.shape and x.dtype == out.dtype
    if x.numel() == 0:
        return
    shape = x.shape
    assert shape[axis] == plan.n, (shape, axis, plan.n)
    outer = int(np.prod(shape[:axis], dtype=np.int64)) if axis > 0 else 1
    inner = int(np.prod(shape[axis + 1 :], dtype=np.int64)) if axis + 1 < len(shape) else 1
    dp = device_plan(plan, x.device)
    with _on_device(x.device):
        fn = getattr(load(), f"fdx_axis_apply_{_sfx(x)}")
        check(
            fn(dp.handle, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), outer, inner, float(alpha), int(bool(accumulate)), _stream_ptr(x.device)),
            "fdx_axis_apply",
        )


def axis_apply_dup(plan: AxisPlan, x: torch.Tensor, outs: Sequence[torch.Tensor], axis: int, alpha: float):
    """outs[0] = outs[1] (= outs[2]) = alpha * D_axis x: one pass that stores its result in up to three arrays (the
    hessian's repeated table entries, reference finite_diff.py:524-529)."""
    outs = list(outs)
    assert 1 <= len(outs) <= 3
    _require(x, "x")
    for o in outs:
        _require(o, "out")
        assert x.shape == o.shape and x.dtype == o.dtype
    if x.numel() == 0:
        return
    shape = x.shape
    assert shape[axis] == plan.n, (shape, axis, plan.n)
    outer = int(np.prod(shape[:axis], dtype=np.int64)) if axis > 0 else 1
    inner = int(np.prod(shape[axis + 1 :], dtype=np.int64)) if axis + 1 < len(shape) else 1
    dp = device_plan(plan, x.device)
    ptr = [C.c_void_p(o.data_ptr()) for o in outs] + [C.c_void_p(0)] * (3 - len(outs))
    with _on_device(x.device):
        fn = getattr(load(), f"fdx_axis_apply_dup_{_sfx(x)}")
        check(fn(dp.handle, C.c_void_p(x.data_ptr()), ptr[0], ptr[1], ptr[2], outer, inner, float(alpha), _stream_ptr(x.device)), "fdx_axis_apply_dup")


def hessian3d(d1: Sequence[AxisPlan], d2: Sequence["AxisPlan | None"], x: torch.Tensor, h: torch.Tensor, a1: Sequence[float],
              a2: Sequence[float]) -> int:
    """The fused 3-D hessian (reference finite_diff.py:460-529): ``h`` (3, 3, *x.shape) receives F[i + j] in every slot,
    one pass over ``x``.  ``d1`` / ``d2``: first- / second-derivative plans per axis (``d2[1]`` unused).  Returns the library
    status: 0, or -2 when the fused kernel does not cover the call (the caller composes axis passes instead)."""
    _require(x, "x")
    _require(h, "h")
    assert x.ndim == 3 and tuple(h.shape) == (3, 3) + tuple(x.shape) and h.dtype == x.dtype
    if x.numel() == 0:
        return 0
    k1 = [device_plan(p, x.device) for p in d1]
    k2 = [device_plan(p, x.device) if p is not None else None for p in d2]
    hp1 = (C.c_void_p * 3)(*[dp.handle for dp in k1])
    hp2 = (C.c_void_p * 3)(*[dp.handle if dp is not None else None for dp in k2])
    with _on_device(x.device):
        fn = getattr(load(), f"fdx_hessian3d_{_sfx(x)}")
        rc = fn(hp1, hp2, C.c_void_p(x.data_ptr()), C.c_void_p(h.data_ptr()), _alphas(a1), _alphas(a2), _stream_ptr(x.device))
    if rc != 0 and rc != -2:
        check(rc, "fdx_hessian3d")
    return rc


def hessian2d(d1: Sequence[AxisPlan], d2: Sequence[AxisPlan], x: torch.Tensor, h: torch.Tensor, a1: Sequence[float], a2: Sequence[float]) -> int:
    """The fused 2-D hessian: ``h`` (2, 2, *x.shape) = [[F0, F1], [F1, F2]] in one pass over ``x``.  Returns 0, or -2 when the
    fused kernel does not cover the call."""
    _require(x, "x")
    _require(h, "h")
    assert x.ndim == 2 and tuple(h.shape) == (2, 2) + tuple(x.shape) and h.dtype == x.dtype
    if x.numel() == 0:
        return 0
    k1 = [device_plan(p, x.device) for p in d1]
    k2 = [device_plan(p, x.device) for p in d2]
    hp1 = (C.c_void_p * 2)(*[dp.handle for dp in k1])
    hp2 = (C.c_void_p * 2)(*[dp.handle for dp in k2])
    al1 = (C.c_double * 2)(*[float(a) for a in a1])
    al2 = (C.c_double * 2)(*[float(a) for a in a2])
    with _on_device(x.device):
        fn = getattr(load(), f"fdx_hessian2d_{_sfx(x)}")
        rc = fn(hp1, hp2, C.c_void_p(x.data_ptr()), C.c_void_p(h.data_ptr()), al1, al2, _stream_ptr(x.device))
    if rc != 0 and rc != -2:
        check(rc, "fdx_hessian2d")
    return rc


def axis_apply_dev_alpha(plan: AxisPlan, x: torch.Tensor, out: torch.Tensor, axis: int, alpha_dev: torch.Tensor, sign: float = 1.0,
                         accumulate: bool = False):
    """out (+)= sign * alpha_dev * D_axis x with the scale read on the device (``alpha_dev``: a float64 CUDA tensor of one
    element, e.g. an element view of a vector of scales).  No host read of the scale: no synchronisation, capture-safe."""
    _require(x, "x")
    _require(out, "out")
    assert x.shape == out.shape and x.dtype == out.dtype
    assert alpha_dev.is_cuda and alpha_dev.dtype == torch.float64 and alpha_dev.numel() >= 1
    if x.numel() == 0:
        return
    shape = x.shape
    assert shape[axis] == plan.n, (shape, axis, plan.n)
    outer = int(np.prod(shape[:axis], dtype=np.int64)) if axis > 0 else 1
    inner = int(np.prod(shape[axis + 1 :], dtype=np.int64)) if axis + 1 < len(shape) else 1
    dp = device_plan(plan, x.device)
    with _on_device(x.device):
        fn = getattr(load(), f"fdx_axis_apply_dev_alpha_{_sfx(x)}")
        check(
            fn(dp.handle, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), outer, inner, C.c_void_p(alpha_dev.data_ptr()), float(sign),
               int(bool(accumulate)), _stream_ptr(x.device)),
            "fdx_axis_apply_dev_alpha",
        )


def fused3d_dev_alpha(kind: str, plans: Sequence[AxisPlan], xs, outs, alpha_dev: torch.Tensor) -> None:
    """The 3-D operators with the three per-axis scales read on the device (``alpha_dev``: float64 CUDA tensor, 3 elements)."""
    x_one, o_one = isinstance(xs, torch.Tensor), isinstance(outs, torch.Tensor)
    first = xs if x_one else xs[0]
    dev = first.device
    assert alpha_dev.is_cuda and alpha_dev.dtype == torch.float64 and alpha_dev.numel() == 3 and alpha_dev.is_contiguous()
    for t in ((xs,) if x_one else tuple(xs)) + ((outs,) if o_one else tuple(outs)):
        _require(t, "tensor")
    keep, hp = _handles(plans, dev)
    fn = getattr(load(), f"fdx_{kind}3d_dev_alpha_{_sfx(first)}")
    xin = C.c_void_p(xs.data_ptr()) if x_one else _ptrs(xs)
    xout = C.c_void_p(outs.data_ptr()) if o_one else _ptrs(outs)
    with _on_device(dev):
        check(fn(hp, xin, xout, C.c_void_p(alpha_dev.data_ptr()), _stream_ptr(dev)), f"fdx_{kind}3d_dev_alpha")
    del keep


def _handles(plans: Sequence[AxisPlan], device):
    dps = [device_plan(p, device) for p in plans]
    arr = (C.c_void_p * 3)(*[dp.handle for dp in dps])
    return dps, arr


def _ptrs(tensors: Sequence[torch.Tensor]):
    return (C.c_void_p * 3)(*[C.c_void_p(t.data_ptr()) for t in tensors])


def _alphas(alpha: Sequence[float]):
    return (C.c_double * 3)(*[float(a) for a in alpha])


def _slab(slab):
    return C.byref(slab) if slab is not None else None


class FusedLauncher:
    """Everything a repeated fused launch needs, built once: entry point, plan-handle array, alpha array and two
    reusable pointer arrays.  ``launch`` only writes the data pointers (computed from the base pointer and the
    component stride: no tensor views) and fetches the raw current stream.  This is the host path of small grids
    (BASELINE config 1: a 16 us kernel behind ~38 us of Python before)."""

    __slots__ = ("kind", "plans", "keep", "hp", "al", "fn", "xin", "xout", "nin", "nout", "dev_index", "x_one", "o_one", "_xp", "_op")

    def __init__(self, kind: str, plans: Sequence[AxisPlan], alpha: Sequence[float], dtype: torch.dtype, device: torch.device):
        if kind not in ("laplacian", "divergence", "gradient", "curl"):
            raise ValueError(kind)
        sfx = "f32" if dtype == torch.float32 else "f64"
        self.kind, self.plans = kind, tuple(plans)
        self.keep, self.hp = _handles(plans, device)
        self.al = _alphas(alpha)
        self.fn = getattr(load(), f"fdx_{kind}3d_{sfx}")
        self.nin = 1 if kind in ("laplacian", "gradient") else 3
        self.nout = 1 if kind in ("laplacian", "divergence") else 3
        self.x_one, self.o_one = self.nin == 1, self.nout == 1
        self.xin = None if self.x_one else (C.c_void_p * 3)()
        self.xout = None if self.o_one else (C.c_void_p * 3)()
        self.dev_index = device.index if device.index is not None else torch.cuda.current_device()

    def launch(self, x_ptr: int, x_stride_bytes: int, out_ptr: int, out_stride_bytes: int) -> int:
        """Enqueue on the current stream of the launcher's device (which must be the current device).  Returns the
        library status: 0, or -2 when the fused kernels do not cover the call."""
        if self.x_one:
            xin = x_ptr
        else:
            xin = self.xin
            xin[0], xin[1], xin[2] = x_ptr, x_ptr + x_stride_bytes, x_ptr + 2 * x_stride_bytes
        if self.o_one:
            xout = out_ptr
        else:
            xout = self.xout
            xout[0], xout[1], xout[2] = out_ptr, out_ptr + out_stride_bytes, out_ptr + 2 * out_stride_bytes
        rc = self.fn(self.hp, xin, xout, self.al, None, torch._C._cuda_getCurrentRawStream(self.dev_index))
        if rc != 0 and rc != -2:
            check(rc, f"fdx_{self.kind}3d")
        return rc


class FusedLauncher2D:
    """Prepared launch of a fused 2-D operator (fdx_*2d_*): ``kind`` in laplacian / divergence / gradient; the field
    indices say which input fields feed the axis-0 / axis-1 stencil and which output fields receive them (the 2-D curl
    is a divergence of (F_1, F_0) with a negated second scale; its adjoint a gradient with swapped outputs)."""

    __slots__ = ("kind", "plans", "keep", "hp", "al", "fn", "fin", "fout", "xin", "xout", "dev_index", "nout")

    def __init__(self, kind: str, plans: Sequence[AxisPlan], alpha: Sequence[float], fin: Sequence[int], fout: Sequence[int], dtype: torch.dtype,
                 device: torch.device):
        sfx = "f32" if dtype == torch.float32 else "f64"
        self.kind, self.plans = kind, tuple(plans)
        self.keep = [device_plan(p, device) for p in plans]
        self.hp = (C.c_void_p * 2)(*[dp.handle for dp in self.keep])
        self.al = (C.c_double * 2)(*[float(a) for a in alpha])
        self.fn = getattr(load(), f"fdx_{kind}2d_{sfx}")
        self.fin, self.fout = tuple(fin), tuple(fout)
        self.xin = (C.c_void_p * 2)() if kind == "divergence" else None
        self.xout = (C.c_void_p * 2)() if kind == "gradient" else None
        self.dev_index = device.index if device.index is not None else torch.cuda.current_device()
        self.nout = len(self.fout)

    def launch(self, x_ptr: int, x_stride_bytes: int, out_ptr: int, out_stride_bytes: int) -> int:
        if self.xin is None:
            xin = x_ptr + self.fin[0] * x_stride_bytes
        else:
            xin = self.xin
            xin[0], xin[1] = x_ptr + self.fin[0] * x_stride_bytes, x_ptr + self.fin[1] * x_stride_bytes
        if self.xout is None:
            xout = out_ptr + self.fout[0] * out_stride_bytes
        else:
            xout = self.xout
            xout[0], xout[1] = out_ptr + self.fout[0] * out_stride_bytes, out_ptr + self.fout[1] * out_stride_bytes
        rc = self.fn(self.hp, xin, xout, self.al, torch._C._cuda_getCurrentRawStream(self.dev_index))
        if rc != 0 and rc != -2:
            check(rc, f"fdx_{self.kind}2d")
        return rc


def fused3d_ptr(kind: str, plans: Sequence[AxisPlan], in_ptrs: Sequence[int], out_ptrs: Sequence[int], alpha: Sequence[float],
                dtype: torch.dtype, device: torch.device, z_begin: int, z_end: int, in_first: int, in_planes: int,
                out_first: int, stream: "torch.cuda.Stream | None" = None) -> int:
    """Raw-pointer form of the fused 3-D entry points for slab execution (multi-GPU halos, host
    streaming).  The plans describe the global grid; ``in_ptrs`` designate global plane ``in_first``
    of buffers holding ``in_planes`` planes, ``out_ptrs`` global plane ``out_first`` (see fdx_slab in
    include/fdx_b200.h).  Returns the library status (0, or -2 = unsupported by the fused kernels)."""
    keep, hp = _handles(plans, device)
    sfx = "f32" if dtype == torch.float32 else "f64"
    st = C.c_void_p((stream or torch.cuda.current_stream(device)).cuda_stream)
    slab = Slab(int(z_begin), int(z_end), int(in_first), int(in_planes), int(out_first))
    lib = load()

    def arr(ptrs):
        ptrs = list(ptrs) + [0] * (3 - len(ptrs))
        return (C.c_void_p * 3)(*[C.c_void_p(int(p)) for p in ptrs])

    with _on_device(device):
        if kind == "laplacian":
            rc = getattr(lib, f"fdx_laplacian3d_{sfx}")(hp, C.c_void_p(int(in_ptrs[0])), C.c_void_p(int(out_ptrs[0])), _alphas(alpha), C.byref(slab), st)
        elif kind == "divergence":
            rc = getattr(lib, f"fdx_divergence3d_{sfx}")(hp, arr(in_ptrs), C.c_void_p(int(out_ptrs[0])), _alphas(alpha), C.byref(slab), st)
        elif kind == "gradient":
            rc = getattr(lib, f"fdx_gradient3d_{sfx}")(hp, C.c_void_p(int(in_ptrs[0])), arr(out_ptrs), _alphas(alpha), C.byref(slab), st)
        elif kind == "curl":
            rc = getattr(lib, f"fdx_curl3d_{sfx}")(hp, arr(in_ptrs), arr(out_ptrs), _alphas(alpha), C.byref(slab), st)
        else:
            raise ValueError(kind)
    del keep
    if rc == -2:
        return rc
    check(rc, f"fdx_{kind}3d (slab)")
    return 0


def prepare_fused3d_ptr(kind: str, plans: Sequence[AxisPlan], in_ptrs: Sequence[int], out_ptrs: Sequence[int], alpha: Sequence[float],
                        dtype: torch.dtype, device: torch.device, z_begin: int, z_end: int, in_first: int, in_planes: int,
                        out_first: int, z_begin2: int = 0, z_end2: int = 0):
    """``fused3d_ptr`` with every ctypes object built once: returns ``launch(stream_ptr: int) -> None`` for callers
    that repeat the same slab call on fixed buffers (the multi-GPU step: three launches per 0.2 ms of GPU work)."""
    keep, hp = _handles(plans, device)
    sfx = "f32" if dtype == torch.float32 else "f64"
    slab = Slab(int(z_begin), int(z_end), int(in_first), int(in_planes), int(out_first), int(z_begin2), int(z_end2))
    fn = getattr(load(), f"fdx_{kind}3d_{sfx}")

    def arr(ptrs):
        ptrs = list(ptrs) + [0] * (3 - len(ptrs))
        return (C.c_void_p * 3)(*[C.c_void_p(int(p)) for p in ptrs])

    xin = C.c_void_p(int(in_ptrs[0])) if kind in ("laplacian", "gradient") else arr(in_ptrs)
    xout = C.c_void_p(int(out_ptrs[0])) if kind in ("laplacian", "divergence") else arr(out_ptrs)
    al, sref = _alphas(alpha), C.byref(slab)

    def launch(stream_ptr: int, _keep=(keep, slab, xin, xout, al)):
        rc = fn(hp, xin, xout, al, sref, C.c_void_p(stream_ptr))
        if rc != 0:
            check(rc, f"fdx_{kind}3d (slab)")

    return launch
