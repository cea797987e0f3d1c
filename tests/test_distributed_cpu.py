"""world_size-2/3 gloo tests (CPU) of the multi-GPU host logic: slab partition, halo exchange
indices, and which input planes a slab call needs.  The CUDA kernels are not involved here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from finitediffx_b200 import distributed as fd
from finitediffx_b200 import slab as fs
from oracle import fdx_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n0, halo, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        full = torch.from_numpy(rng.standard_normal((n0, 5, 6)))
        sl = fd.Slab(global_n0=n0, rank=rank, world=world)
        owned = sl.z1 - sl.z0
        buf = torch.zeros((owned + 2 * halo, 5, 6), dtype=torch.float64)
        buf[halo : halo + owned] = full[sl.z0 : sl.z1]
        for r in fd.exchange_halos([buf], owned, halo, sl):
            r.wait()
        # every plane of the buffer that exists globally must now hold the global data
        ok = True
        for i in range(owned + 2 * halo):
            z = sl.z0 - halo + i
            inside = 0 <= z < n0
            neighbour_side = (i < halo and sl.lower is not None) or (i >= halo + owned and sl.upper is not None) or (halo <= i < halo + owned)
            if inside and neighbour_side:
                ok &= bool(torch.equal(buf[i], full[z]))
        # the slab, differentiated along axis 0 with the reference's stencils through the oracle on
        # the padded local buffer, must reproduce the global result on the planes whose stencils see
        # only interior points of the padded buffer
        ref = O.difference(full.numpy(), axis=0, accuracy=2 * halo - 1 if halo > 1 else 1, derivative=1)
        acc = 2 * halo - 1 if halo > 1 else 1
        lo_pad = halo if sl.lower is not None else 0
        hi_pad = halo if sl.upper is not None else 0
        local = buf[halo - lo_pad : halo + owned + hi_pad].numpy()
        got = O.difference(local, axis=0, accuracy=acc, derivative=1)[lo_pad : lo_pad + owned]
        # interior ranks: central rows everywhere; end ranks: one-sided rows coincide with the global ones
        ok &= bool(np.allclose(got, ref[sl.z0 : sl.z1], rtol=1e-12, atol=1e-12))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n0,halo", [(2, 24, 2), (3, 31, 3), (2, 17, 1)])
def test_halo_exchange_and_partition_gloo(world, n0, halo):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n0, halo, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert all(ret[r] for r in range(world)), dict(ret)


def test_slab_partition_covers_the_axis():
    for n0, world in ((512, 8), (100, 3), (17, 4), (2048, 8)):
        cover = []
        for r in range(world):
            sl = fd.Slab(n0, r, world)
            cover += list(range(sl.z0, sl.z1))
            assert (sl.lower is None) == (r == 0)
            if sl.upper is None:
                assert sl.z1 == n0
        assert cover == list(range(n0))


def test_input_planes_of_a_slab_call():
    spec = fs.make_spec("laplacian", (64, 32, 32), accuracy=4, step_size=0.1)  # P=3, band rows 3, band width 8
    assert spec.radius == 3
    assert spec.input_planes(16, 32) == (13, 35)
    assert spec.input_planes(0, 16) == (0, 19)      # lower boundary: planes under the one-sided stencils
    assert spec.input_planes(0, 2) == (0, 8)
    assert spec.input_planes(60, 64) == (56, 64)
    assert spec.input_planes(3, 10) == (0, 13)
    spec = fs.make_spec("divergence", (64, 32, 32), accuracy=2, step_size=(0.1, 0.2, 0.3))
    assert spec.radius == 1 and spec.alpha == (10.0, 5.0, 1.0 / 0.3)
    with pytest.raises(ValueError):
        fs.make_spec("hessian", (64, 32, 32), accuracy=2, step_size=0.1)


def test_thin_end_slabs_are_rejected():
    """The interior launch of an end rank holds the planes under the one-sided boundary stencils and does not wait
    for the halo exchange, so those stencils must read owned planes only: owned >= band width (P = 3: 8 planes).
    owned = 6, 7 used to pass the check and race with the incoming halo."""
    cpu = torch.device("cpu")
    for owned, ok in ((6, False), (7, False), (8, True), (16, True)):
        for rank in (0, 3):
            sl = fd.Slab(global_n0=4 * owned, rank=rank, world=4)
            if ok:
                op = fd.SlabOperator(sl, "laplacian", (4 * owned, 16, 16), torch.float32, accuracy=4, step_size=0.1, device=cpu)
                assert op.owned == owned and op.P == 3
            else:
                with pytest.raises(ValueError):
                    fd.SlabOperator(sl, "laplacian", (4 * owned, 16, 16), torch.float32, accuracy=4, step_size=0.1, device=cpu)
    # interior ranks only need 2P planes
    fd.SlabOperator(fd.Slab(24, 1, 4), "laplacian", (24, 16, 16), torch.float32, accuracy=4, step_size=0.1, device=cpu)
