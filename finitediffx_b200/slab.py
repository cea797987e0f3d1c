"""Slab execution of the fused 3-D operators along axis 0.

One mechanism serves two callers:

* ``HostStreamer`` -- grids that live in (pinned) host memory, possibly larger than HBM: slabs of
  planes are copied in, differentiated and copied out on three CUDA streams so that the H2D copy,
  the kernel and the D2H copy of neighbouring slabs overlap.  This is what the public operators
  use when they are handed host arrays, and what bench.py's ``e2e`` number measures.
* ``finitediffx_b200.distributed.SlabOperator`` -- one slab per GPU with NCCL halo exchange.

A slab call tells the kernel which global plane its buffers start at and which output planes to
produce (``fdx_slab`` in include/fdx_b200.h); only planes ``[z_begin - P, z_end + P)`` (clipped to the
grid) and, next to a physical boundary, the planes under the one-sided stencils are touched, and
those are what the device buffers hold.  Boundary rows therefore use exactly the reference's stencils
(finitediffx/_src/finite_diff.py:75-88) no matter how the grid is cut.
"""
from __future__ import annotations

import dataclasses as dc
from typing import Sequence

import torch

from . import _cabi
from .plan import AxisPlan, axis_plan

# bytes of one field slab the host streamer moves per pipeline step (H2D / kernel / D2H overlap)
DEFAULT_SLAB_BYTES = 64 << 20

FIELDS = {"laplacian": (1, 1), "divergence": (3, 1), "gradient": (1, 3), "curl": (3, 3)}


@dc.dataclass(frozen=True)
class SlabSpec:
    """What a fused operator on a global (n0, n1, n2) grid needs."""

    kind: str
    spatial: tuple
    plans: tuple  # three AxisPlans (global axis lengths)
    alpha: tuple  # three floats

    @property
    def radius(self) -> int:
        return self.plans[0].hi

    def input_planes(self, z_begin: int, z_end: int) -> tuple:
        """Planes of the input the kernel reads to produce output planes [z_begin, z_end)."""
        p0, n0 = self.plans[0], self.spatial[0]
        lo, hi = max(0, z_begin + p0.lo), min(n0, z_end + p0.hi)
        if z_begin < p0.nl:  # output planes under the lower one-sided stencils
            lo, hi = 0, max(hi, p0.wl)
        if z_end > n0 - p0.nr:
            lo, hi = min(lo, n0 - p0.wr), n0
        return lo, hi


def make_spec(kind: str, spatial: Sequence[int], *, accuracy: int, step_size, method: str = "central", derivative=None) -> SlabSpec:
    if kind not in FIELDS:
        raise ValueError(kind)
    d = derivative if derivative is not None else (2 if kind == "laplacian" else 1)
    steps = step_size if isinstance(step_size, (tuple, list)) else (step_size,) * 3
    plans = tuple(axis_plan(int(spatial[k]), d, int(accuracy), method, k) for k in range(3))
    alpha = tuple(1.0 / float(h) ** d for h in steps)
    return SlabSpec(kind, tuple(int(s) for s in spatial), plans, alpha)


def run_slab(spec: SlabSpec, in_bufs: Sequence[torch.Tensor], in_first_plane: int, out_bufs: Sequence[torch.Tensor],
             out_first_plane: int, z_begin: int, z_end: int, stream: "torch.cuda.Stream | None" = None) -> None:
    """Produce output planes [z_begin, z_end).  ``in_bufs[f]`` holds input planes starting at global
    plane ``in_first_plane`` of field f (shape (planes, n1, n2)); ``out_bufs[o]`` receives output
    planes starting at global plane ``out_first_plane``."""
    n0, n1, n2 = spec.spatial
    lo, hi = spec.input_planes(z_begin, z_end)
    for b in in_bufs:
        assert b.is_cuda and b.is_contiguous() and b.shape[1:] == (n1, n2)
        assert in_first_plane <= lo and hi <= in_first_plane + b.shape[0], (in_first_plane, b.shape[0], lo, hi)
    for b in out_bufs:
        assert b.is_cuda and b.is_contiguous() and b.shape[1:] == (n1, n2)
        assert out_first_plane <= z_begin and z_end <= out_first_plane + b.shape[0]
    rc = _cabi.fused3d_ptr(spec.kind, spec.plans, [b.data_ptr() for b in in_bufs], [b.data_ptr() for b in out_bufs], spec.alpha,
                           in_bufs[0].dtype, in_bufs[0].device, z_begin, z_end, in_first_plane, in_bufs[0].shape[0],
                           out_first_plane, stream=stream)
    if rc != 0:
        raise _cabi.FdxError(
            f"the fused {spec.kind} kernel does not cover this slab call (grid {spec.spatial}, radius {spec.radius}); "
            "slab execution has no composed fallback"
        )


class HostStreamer:
    """Out-of-core execution: host (pinned) arrays in, host arrays out, slabs pipelined over three
    CUDA streams.  Device memory use is ``n_buffers`` slabs regardless of the grid size."""

    def __init__(self, spec: SlabSpec, dtype: torch.dtype, device: torch.device, slab_planes: int = 64, n_buffers: int = 3):
        self.spec, self.dtype, self.device = spec, dtype, device
        n0, n1, n2 = spec.spatial
        p0 = spec.plans[0]
        self.m = max(min(slab_planes, n0), p0.wl, p0.wr, 1)
        self.nb = n_buffers
        nin, nout = FIELDS[spec.kind]
        reach = max(-p0.lo, p0.hi)
        depth = self.m + 2 * reach + max(p0.wl, p0.wr)  # generous: band planes ride along at the ends
        self.din = [[torch.empty((depth, n1, n2), dtype=dtype, device=device) for _ in range(nin)] for _ in range(n_buffers)]
        self.dout = [[torch.empty((self.m, n1, n2), dtype=dtype, device=device) for _ in range(nout)] for _ in range(n_buffers)]
        self.s_in, self.s_c, self.s_out = (torch.cuda.Stream(device) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(n_buffers)]
        self.ev_c = [torch.cuda.Event() for _ in range(n_buffers)]
        self.ev_out = [torch.cuda.Event() for _ in range(n_buffers)]
        self.bytes_h2d = self.bytes_d2h = 0

    def run(self, x_host: Sequence[torch.Tensor], out_host: Sequence[torch.Tensor]) -> None:
        spec = self.spec
        n0 = spec.spatial[0]
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_c, self.s_out):
            s.wait_stream(cur)
        k = 0
        self.bytes_h2d = self.bytes_d2h = 0
        for z0 in range(0, n0, self.m):
            z1 = min(n0, z0 + self.m)
            b = k % self.nb
            lo, hi = spec.input_planes(z0, z1)
            with torch.cuda.stream(self.s_in):
                if k >= self.nb:
                    self.s_in.wait_event(self.ev_c[b])  # the kernel that read this input buffer is done
                for f, xh in enumerate(x_host):
                    self.din[b][f][: hi - lo].copy_(xh[lo:hi], non_blocking=True)
                    self.bytes_h2d += (hi - lo) * xh[0].numel() * xh.element_size()
                self.ev_in[b].record(self.s_in)
            with torch.cuda.stream(self.s_c):
                self.s_c.wait_event(self.ev_in[b])
                if k >= self.nb:
                    self.s_c.wait_event(self.ev_out[b])  # the copy-out of this output buffer is done
                run_slab(spec, [t[: hi - lo] for t in self.din[b]], lo, self.dout[b], z0, z0, z1, stream=self.s_c)
                self.ev_c[b].record(self.s_c)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_c[b])
                for o, oh in enumerate(out_host):
                    oh[z0:z1].copy_(self.dout[b][o][: z1 - z0], non_blocking=True)
                    self.bytes_d2h += (z1 - z0) * oh[0].numel() * oh.element_size()
                self.ev_out[b].record(self.s_out)
            k += 1
        for s in (self.s_in, self.s_c, self.s_out):
            cur.wait_stream(s)
