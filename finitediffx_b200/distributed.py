"""Multi-GPU slab decomposition with halo exchange (one process per GPU, torch.distributed/NCCL).

The reference has no distributed path; this is the new driver the north star
asks for: a large 3-D grid is cut into slabs along axis 0, rank r owning planes [z0, z1).  Only the
derivative along axis 0 crosses a cut, so each operator call needs ONE neighbour exchange of
``P`` planes (P = stencil radius) per side, for the fields that are differentiated along axis 0:

    laplacian / gradient : the scalar field
    divergence           : component 0 only
    curl                 : components 1 and 2 only

Halo planes are contiguous in a C-ordered array, so they are sent in place (no packing) with NCCL
point-to-point send/recv between neighbours, on a communication stream.  The fused kernel for the
planes that do not depend on a halo ([z0+P, z1-P)) runs concurrently on the compute stream; the two
P-plane shells next to the cuts are launched on a third stream behind the halo event, so they fill
the tail of the interior kernel.  Ranks at the physical ends of
the grid keep the reference's one-sided boundary stencils: every rank addresses planes by their
GLOBAL index (finitediffx_b200/slab.py), so the boundary logic is the single-GPU one.
"""
from __future__ import annotations

import dataclasses as dc
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import slab as _slab

# fields (input components) whose derivative along axis 0 is taken, i.e. whose halos must be real
Z_FIELDS = {"laplacian": (0,), "gradient": (0,), "divergence": (0,), "curl": (1, 2)}


@dc.dataclass(frozen=True)
class Slab:
    """Rank r of `world` owns global planes [z0, z1) of an axis of length global_n0."""

    global_n0: int
    rank: int
    world: int
    group: Optional[object] = None

    @property
    def planes_per_rank(self) -> int:
        return -(-self.global_n0 // self.world)

    @property
    def z0(self) -> int:
        return min(self.global_n0, self.rank * self.planes_per_rank)

    @property
    def z1(self) -> int:
        return min(self.global_n0, self.z0 + self.planes_per_rank)

    @property
    def lower(self) -> Optional[int]:
        return self.rank - 1 if self.rank > 0 else None

    @property
    def upper(self) -> Optional[int]:
        return self.rank + 1 if self.rank < self.world - 1 and self.z1 < self.global_n0 else None


def exchange_halos(bufs: Sequence[torch.Tensor], owned: int, halo: int, sl: Slab) -> List:
    """Neighbour halo exchange for buffers laid out [halo | owned planes | halo] along dim 0.

    Sends the first/last ``halo`` owned planes to the lower/upper neighbour and receives their
    counterparts into the halo regions.  Works on any backend (NCCL on GPUs, gloo in the CPU tests).
    Returns the list of outstanding requests (call ``.wait()`` on each)."""
    ops = []
    for b in bufs:
        assert b.shape[0] == owned + 2 * halo and b.is_contiguous()
        if sl.lower is not None:
            ops.append(dist.P2POp(dist.isend, b[halo : 2 * halo], sl.lower, sl.group))
            ops.append(dist.P2POp(dist.irecv, b[0:halo], sl.lower, sl.group))
        if sl.upper is not None:
            ops.append(dist.P2POp(dist.isend, b[owned : owned + halo], sl.upper, sl.group))
            ops.append(dist.P2POp(dist.irecv, b[owned + halo : owned + 2 * halo], sl.upper, sl.group))
    if not ops or halo == 0:
        return []
    return dist.batch_isend_irecv(ops)


class SlabOperator:
    """A fused operator on a slab-decomposed global grid.

    Device buffers (per rotating buffer set): inputs [P halo | owned | P halo] planes, outputs
    [owned] planes.  ``step(b)`` = halo exchange (comm stream) overlapped with the interior kernel;
    the two shells run on a third stream as soon as the halos have landed.  Requires owned >= 2P
    and, on the end ranks, owned >= the width of the one-sided boundary band (its planes are read by the
    interior launch, which does not wait for the halos)."""

    def __init__(self, sl: Slab, kind: str, spatial: Sequence[int], dtype: torch.dtype, *, accuracy: int, step_size,
                 device: torch.device, method: str = "central", n_buffers: int = 1):
        n1, n2 = int(spatial[-2]), int(spatial[-1])
        self.sl, self.kind, self.dtype, self.device = sl, kind, dtype, device
        self.spec = _slab.make_spec(kind, (sl.global_n0, n1, n2), accuracy=accuracy, step_size=step_size, method=method)
        p0 = self.spec.plans[0]
        self.P = max(-p0.lo, p0.hi)
        self.owned = sl.z1 - sl.z0
        # The planes under the one-sided boundary stencils belong to the interior launch, which does not wait for the
        # halo exchange: on the end ranks every plane those stencils read (wl / wr planes from the physical boundary)
        # must be an OWNED plane, never a halo plane that NCCL may still be writing.
        if self.owned < 2 * self.P or (sl.lower is None and sl.upper is not None and self.owned < p0.wl) or (
                sl.upper is None and sl.lower is not None and self.owned < p0.wr):
            raise ValueError(f"slab of {self.owned} planes is too thin for stencil radius {self.P} / boundary bands {p0.wl},{p0.wr}")
        nin, nout = _slab.FIELDS[kind]
        depth = self.owned + 2 * self.P
        self.inputs = [[torch.zeros((depth, n1, n2), dtype=dtype, device=device) for _ in range(nin)] for _ in range(n_buffers)]
        self.outputs = [[torch.empty((self.owned, n1, n2), dtype=dtype, device=device) for _ in range(nout)] for _ in range(n_buffers)]
        # high priority: the exchange and the shells must not queue behind the thousands of CTAs of the interior kernel
        self.comm = torch.cuda.Stream(device, priority=-1) if device.type == "cuda" else None
        self.ev_halo = torch.cuda.Event() if device.type == "cuda" else None
        # the two shells run on their own stream: they become runnable as soon as the halos have landed
        # and fill the tail of the interior kernel instead of following it
        self.shell = torch.cuda.Stream(device, priority=-1) if device.type == "cuda" else None
        self.ev_shell = torch.cuda.Event() if device.type == "cuda" else None
        zf_n = len(Z_FIELDS[kind])
        # bytes this rank receives (= sends) per step
        self.halo_bytes_per_step = zf_n * self.P * n1 * n2 * torch.empty((), dtype=dtype).element_size() * ((sl.lower is not None) + (sl.upper is not None))
        self.graphs = {}
        self._prep = {}
        self.launches_per_step = 0

    # -- CUDA graphs: a step is ~10 enqueue calls (NCCL group, events, three launches); at 0.2 ms of GPU work
    # per step the host is the bottleneck, so a step on fixed buffers can be captured once and replayed
    def capture(self, b: int = 0) -> None:
        """Capture ``step(b)`` (halo exchange on the comm stream, interior kernel, shells) in a CUDA graph."""
        from . import _cabi

        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.step(b)  # warm-up on the capture stream
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        n0 = _cabi.launch_count()
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
            self.step(b)
        self.launches_per_step = _cabi.launch_count() - n0
        self.graphs[b] = g

    def replay(self, b: int = 0) -> List[torch.Tensor]:
        self.graphs[b].replay()
        return self.outputs[b]

    # -- data access -----------------------------------------------------------------------
    def owned_view(self, b: int, f: int = 0) -> torch.Tensor:
        """The planes this rank owns inside input buffer b, field f (write the field here)."""
        return self.inputs[b][f][self.P : self.P + self.owned]

    def fill_random(self, b: int, seed: int) -> None:
        g = torch.Generator(device=self.device).manual_seed(seed)
        for f in range(len(self.inputs[b])):
            v = self.owned_view(b, f)
            for z in range(0, v.shape[0], 64):  # in pieces: a 2048^3 slab must not need a second copy of itself
                v[z : z + 64].normal_(generator=g)

    # -- one operator application ------------------------------------------------------------
    def _prepared(self, b: int):
        """Per rotating buffer: the P2P ops of the exchange and the three prepared kernel launches (the
        buffers are fixed, so none of this is rebuilt per step: the step is host-bound otherwise)."""
        hit = self._prep.get(b)
        if hit is not None:
            return hit
        from . import _cabi

        sl, P = self.sl, self.P
        z0, z1 = sl.z0, sl.z1
        ins, outs = self.inputs[b], self.outputs[b]
        ops = []
        for t in (ins[f] for f in Z_FIELDS[self.kind]):
            if sl.lower is not None:
                ops.append(dist.P2POp(dist.isend, t[P : 2 * P], sl.lower, sl.group))
                ops.append(dist.P2POp(dist.irecv, t[0:P], sl.lower, sl.group))
            if sl.upper is not None:
                ops.append(dist.P2POp(dist.isend, t[self.owned : self.owned + P], sl.upper, sl.group))
                ops.append(dist.P2POp(dist.irecv, t[self.owned + P : self.owned + 2 * P], sl.upper, sl.group))
        lo = z0 + (P if sl.lower is not None else 0)
        hi = z1 - (P if sl.upper is not None else 0)

        def prep(za, zb):
            lo_in, hi_in = self.spec.input_planes(za, zb)
            assert z0 - P <= lo_in and hi_in <= z0 - P + ins[0].shape[0], (za, zb, lo_in, hi_in)
            return _cabi.prepare_fused3d_ptr(self.kind, self.spec.plans, [t.data_ptr() for t in ins], [t.data_ptr() for t in outs],
                                             self.spec.alpha, self.dtype, self.device, za, zb, z0 - P, ins[0].shape[0], z0)

        interior = prep(lo, hi)
        shells = []
        if sl.lower is not None:
            shells.append(prep(z0, lo))
        if sl.upper is not None:
            shells.append(prep(hi, z1))
        hit = (ops, interior, shells, prep(z0, z1))
        self._prep[b] = hit
        return hit

    def step_alone(self, b: int = 0) -> List[torch.Tensor]:
        """The same output planes in ONE launch with no exchange (the halo planes keep what the last exchange left):
        the time of this rank's share of the kernel work alone, the reference point of the overlap efficiency."""
        self._prepared(b)[3](torch.cuda.current_stream(self.device).cuda_stream)
        return self.outputs[b]

    def step(self, b: int = 0) -> List[torch.Tensor]:
        sl = self.sl
        cur = torch.cuda.current_stream(self.device)
        outs = self.outputs[b]
        if self.device.type != "cuda":
            raise RuntimeError("SlabOperator.step needs CUDA buffers (there is no CPU path)")
        ops, interior, shells, _ = self._prepared(b)
        has_nb = bool(shells)
        if has_nb:
            self.comm.wait_stream(cur)  # inputs of this step are ready
            with torch.cuda.stream(self.comm):
                if ops and self.P > 0:
                    for r in dist.batch_isend_irecv(ops):
                        r.wait()
                self.ev_halo.record(self.comm)
            self.shell.wait_stream(cur)  # (before the interior launch: only the inputs of this step)
        # planes that need no halo: everything except P planes next to each cut
        interior(cur.cuda_stream)
        if has_nb:
            self.shell.wait_event(self.ev_halo)
            sp = self.shell.cuda_stream
            for launch in shells:
                launch(sp)
            self.ev_shell.record(self.shell)
            cur.wait_event(self.ev_shell)
        return outs
