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

    Device buffers (per rotating buffer set): inputs [P halo | owned | P halo] planes, outputs [owned] planes.
    ``step(b)`` = halo exchange (comm stream) overlapped with the interior kernel; ONE launch for the two P-plane shells
    follows the exchange on the same stream.  Requires owned >= 2P and, on the end ranks, owned >= the width of the
    one-sided boundary band (its planes are read by the interior launch, which does not wait for the halos).

    Halo transport (``halo=`` or FDX_HALO):

    * ``"peer"`` -- the buffers of the fields that are differentiated along axis 0 live in torch symmetric memory
      (CUDA backend: every rank maps its neighbours' buffers over NVLink).  A rank PULLS its halo planes with
      copy-engine peer copies between two signal-pad barriers: no NCCL kernels on the SMs, nothing but memcpys and
      kernels in the step, so the step is captured in a CUDA graph and replayed as one launch.
    * ``"nccl"`` -- ``batch_isend_irecv`` of the halo planes (the round-1 path; the fallback when symmetric memory
      cannot be set up).
    """

    def __init__(self, sl: Slab, kind: str, spatial: Sequence[int], dtype: torch.dtype, *, accuracy: int, step_size,
                 device: torch.device, method: str = "central", n_buffers: int = 1, halo: Optional[str] = None):
        import os

        n1, n2 = int(spatial[-2]), int(spatial[-1])
        self.sl, self.kind, self.dtype, self.device = sl, kind, dtype, device
        self.spec = _slab.make_spec(kind, (sl.global_n0, n1, n2), accuracy=accuracy, step_size=step_size, method=method)
        p0 = self.spec.plans[0]
        self.P = max(-p0.lo, p0.hi)
        self.owned = sl.z1 - sl.z0
        # The planes under the one-sided boundary stencils belong to the interior launch, which does not wait for the
        # halo exchange: on the end ranks every plane those stencils read (wl / wr planes from the physical boundary)
        # must be an OWNED plane, never a halo plane that the exchange may still be writing.
        if self.owned < 2 * self.P or (sl.lower is None and sl.upper is not None and self.owned < p0.wl) or (
                sl.upper is None and sl.lower is not None and self.owned < p0.wr):
            raise ValueError(f"slab of {self.owned} planes is too thin for stencil radius {self.P} / boundary bands {p0.wl},{p0.wr}")
        nin, nout = _slab.FIELDS[kind]
        depth = self.owned + 2 * self.P
        self.depth = depth
        cuda = device.type == "cuda"
        want = (halo or os.environ.get("FDX_HALO") or "peer").lower()
        self.mode = "nccl"
        self._symm = None
        zf = Z_FIELDS[kind]
        self.inputs = None
        if cuda and want == "peer" and sl.world > 1 and dist.is_available() and dist.is_initialized():
            try:
                self._setup_peer(n_buffers, nin, depth, n1, n2)
                self.mode = "peer"
            except Exception as e:  # symmetric memory is optional infrastructure: NCCL send/recv is the spec'd path
                self._symm = None
                self.inputs = None
                self.peer_error = repr(e)[:200]
        if self.inputs is None:
            self.inputs = [[torch.zeros((depth, n1, n2), dtype=dtype, device=device) for _ in range(nin)] for _ in range(n_buffers)]
        self.outputs = [[torch.empty((self.owned, n1, n2), dtype=dtype, device=device) for _ in range(nout)] for _ in range(n_buffers)]
        # high priority: the exchange and the shells must not queue behind the thousands of CTAs of the interior kernel
        self.comm = torch.cuda.Stream(device, priority=-1) if cuda else None
        self.ev_done = torch.cuda.Event() if cuda else None
        # bytes this rank receives (= sends) per step
        self.halo_bytes_per_step = len(zf) * self.P * n1 * n2 * torch.empty((), dtype=dtype).element_size() * ((sl.lower is not None) + (sl.upper is not None))
        self.graphs = {}
        self._prep = {}
        self.launches_per_step = 0

    # -- symmetric memory: the z-field buffers of all rotating sets in ONE allocation, one rendezvous
    def _setup_peer(self, n_buffers: int, nin: int, depth: int, n1: int, n2: int) -> None:
        import torch.distributed._symmetric_memory as symm

        zf = Z_FIELDS[self.kind]
        group = self.sl.group if self.sl.group is not None else dist.group.WORLD
        if dist.get_world_size(group) != self.sl.world or dist.get_rank(group) != self.sl.rank:
            raise RuntimeError("slab rank order must be the group's rank order")
        elems = n_buffers * len(zf) * depth * n1 * n2
        pool = symm.empty(elems, dtype=self.dtype, device=self.device)
        pool.zero_()
        hdl = symm.rendezvous(pool, group)
        self._symm, self._pool = hdl, pool
        plane = n1 * n2
        self._peer_off = lambda b, i: (b * len(zf) + i) * depth * plane  # element offset of (buffer set b, z-field i)
        self.inputs = []
        for b in range(n_buffers):
            row = []
            for f in range(nin):
                if f in zf:
                    o = self._peer_off(b, zf.index(f))
                    row.append(pool[o : o + depth * plane].view(depth, n1, n2))
                else:
                    row.append(torch.zeros((depth, n1, n2), dtype=self.dtype, device=self.device))
            self.inputs.append(row)
        self._peer_views = {}
        for nb in (self.sl.lower, self.sl.upper):
            if nb is not None:
                self._peer_views[nb] = hdl.get_buffer(nb, (elems,), self.dtype)
        hdl.barrier(channel=0)

    # -- CUDA graphs: a step is a dozen enqueue calls (barriers, copies, events, launches); at 0.2 ms of GPU work
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

    def release_graphs(self) -> None:
        """Drop captured steps (before the process group is torn down)."""
        torch.cuda.synchronize(self.device)
        self.graphs.clear()

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
        """Per rotating buffer: the exchange (P2P ops or peer copies) and the prepared kernel launches (the buffers are
        fixed, so none of this is rebuilt per step: the step is host-bound otherwise)."""
        hit = self._prep.get(b)
        if hit is not None:
            return hit
        from . import _cabi

        sl, P, owned = self.sl, self.P, self.owned
        z0, z1 = sl.z0, sl.z1
        ins, outs = self.inputs[b], self.outputs[b]
        zf = Z_FIELDS[self.kind]
        ops, copies = [], []
        if self.mode == "peer":
            plane = ins[0].shape[1] * ins[0].shape[2]
            for i, f in enumerate(zf):
                t = ins[f]
                o = self._peer_off(b, i)
                if sl.lower is not None:  # my lower halo <- the lower neighbour's last P owned planes
                    src = self._peer_views[sl.lower][o + owned * plane : o + (owned + P) * plane].view(P, *t.shape[1:])
                    copies.append((t[0:P], src))
                if sl.upper is not None:  # my upper halo <- the upper neighbour's first P owned planes
                    src = self._peer_views[sl.upper][o + P * plane : o + 2 * P * plane].view(P, *t.shape[1:])
                    copies.append((t[owned + P : owned + 2 * P], src))
        else:
            for t in (ins[f] for f in zf):
                if sl.lower is not None:
                    ops.append(dist.P2POp(dist.isend, t[P : 2 * P], sl.lower, sl.group))
                    ops.append(dist.P2POp(dist.irecv, t[0:P], sl.lower, sl.group))
                if sl.upper is not None:
                    ops.append(dist.P2POp(dist.isend, t[owned : owned + P], sl.upper, sl.group))
                    ops.append(dist.P2POp(dist.irecv, t[owned + P : owned + 2 * P], sl.upper, sl.group))
        lo = z0 + (P if sl.lower is not None else 0)
        hi = z1 - (P if sl.upper is not None else 0)

        def prep(za, zb, za2=0, zb2=0):
            for a_, b_ in ((za, zb), (za2, zb2)):
                if b_ > a_:
                    lo_in, hi_in = self.spec.input_planes(a_, b_)
                    assert z0 - P <= lo_in and hi_in <= z0 - P + ins[0].shape[0], (a_, b_, lo_in, hi_in)
            return _cabi.prepare_fused3d_ptr(self.kind, self.spec.plans, [t.data_ptr() for t in ins], [t.data_ptr() for t in outs],
                                             self.spec.alpha, self.dtype, self.device, za, zb, z0 - P, ins[0].shape[0], z0, za2, zb2)

        interior = prep(lo, hi)
        ranges = [r for r in ((z0, lo), (hi, z1)) if r[1] > r[0]]
        shell = None
        if len(ranges) == 2:  # both shells in ONE launch (second range of fdx_slab)
            shell = prep(ranges[0][0], ranges[0][1], ranges[1][0], ranges[1][1])
        elif ranges:
            shell = prep(ranges[0][0], ranges[0][1])
        hit = (ops, copies, interior, shell, prep(z0, z1))
        self._prep[b] = hit
        return hit

    def step_alone(self, b: int = 0) -> List[torch.Tensor]:
        """The same output planes in ONE launch with no exchange (the halo planes keep what the last exchange left):
        the time of this rank's share of the kernel work alone, the reference point of the overlap efficiency."""
        self._prepared(b)[4](torch.cuda.current_stream(self.device).cuda_stream)
        return self.outputs[b]

    def step(self, b: int = 0, timeline: Optional[dict] = None) -> List[torch.Tensor]:
        """One operator application.  ``timeline``: a dict that receives CUDA events (start, halo, shell, interior,
        done) of this step -- the record kept under profiles/ comes from these (see tools/dist_timeline.py)."""
        cur = torch.cuda.current_stream(self.device)
        outs = self.outputs[b]
        if self.device.type != "cuda":
            raise RuntimeError("SlabOperator.step needs CUDA buffers (there is no CPU path)")
        ops, copies, interior, shell, _ = self._prepared(b)

        def mark(name, stream):
            if timeline is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                timeline[name] = ev

        mark("start", cur)
        if shell is not None:
            self.comm.wait_stream(cur)  # inputs of this step are ready
            with torch.cuda.stream(self.comm):
                if self.mode == "peer":
                    self._symm.barrier(channel=0)  # every rank's inputs of this step are in place
                    for dst, src in copies:
                        dst.copy_(src, non_blocking=True)  # copy engine over NVLink: no SM is taken from the interior kernel
                elif ops and self.P > 0:
                    for r in dist.batch_isend_irecv(ops):
                        r.wait()
                mark("halo", self.comm)
                shell(self.comm.cuda_stream)  # behind the halos, same stream: no event hop
                mark("shell", self.comm)
                if self.mode == "peer":
                    self._symm.barrier(channel=1)  # every rank has pulled: the inputs may be overwritten after this step
                self.ev_done.record(self.comm)
        # planes that need no halo: everything except P planes next to each cut
        interior(cur.cuda_stream)
        mark("interior", cur)
        if shell is not None:
            cur.wait_event(self.ev_done)
        mark("done", cur)
        return outs
