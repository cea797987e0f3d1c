"""GPU tests of slab execution: host streaming (out-of-core) and the multi-GPU slab driver emulated
on one GPU (each "rank" is a separate set of buffers; halos are copied by hand instead of NCCL)."""
import numpy as np
import pytest
import torch

import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi
from finitediffx_b200 import slab as fs
from oracle import fdx_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["laplacian", "divergence", "gradient", "curl"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_slabs_reproduce_the_whole_grid(kind, dtype):
    """Cut axis 0 into 3 uneven slabs with P-plane halos; every slab call must give exactly the
    planes of the single-call result (same kernels, same taps, boundary rows included)."""
    torch.manual_seed(0)
    n0, n1, n2 = 70, 40, 64
    acc = 4
    nin, nout = fs.FIELDS[kind]
    x = torch.randn((nin, n0, n1, n2), device="cuda", dtype=dtype)
    steps = (0.1, 0.2, 0.3)
    spec = fs.make_spec(kind, (n0, n1, n2), accuracy=acc, step_size=steps)
    P = spec.radius
    full = [torch.empty((n0, n1, n2), device="cuda", dtype=dtype) for _ in range(nout)]
    fs.run_slab(spec, [x[f] for f in range(nin)], 0, full, 0, 0, n0)
    cuts = [0, 23, 41, n0]
    for z0, z1 in zip(cuts[:-1], cuts[1:]):
        lo, hi = spec.input_planes(z0, z1)
        ins = [x[f][lo:hi].clone() for f in range(nin)]  # this "rank" only holds its planes + halos
        outs = [torch.full((z1 - z0, n1, n2), float("nan"), device="cuda", dtype=dtype) for _ in range(nout)]
        # interior first, shells after -- as the distributed driver does
        mid0, mid1 = min(z0 + P, z1), max(z1 - P, min(z0 + P, z1))
        for a, b in ((mid0, mid1), (z0, mid0), (mid1, z1)):
            if b > a:
                fs.run_slab(spec, ins, lo, outs, z0, a, b)
        for o in range(nout):
            assert torch.equal(outs[o], full[o][z0:z1]), (kind, z0, z1, o)
    # and the whole-grid slab call equals the public operator
    xin = x[0] if nin == 1 else x
    kw = dict(accuracy=acc, step_size=steps)
    ref = fdx.divergence(xin, keepdims=False, **kw) if kind == "divergence" else getattr(fdx, kind)(xin, **kw)
    got = full[0] if nout == 1 else torch.stack(full)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("kind", ["laplacian", "divergence", "gradient", "curl"])
def test_host_arrays_stream_through_the_gpu(kind, monkeypatch):
    """Host arrays >= 2^24 elements take the pipelined slab path; results must equal the device path."""
    rng = np.random.default_rng(1)
    monkeypatch.setattr(fs, "DEFAULT_SLAB_BYTES", 8 << 20)  # several slabs even on this small grid
    nin, nout = fs.FIELDS[kind]
    shape = (256, 256, 256) if nin == 1 else (3, 192, 192, 160)
    x = rng.standard_normal(shape).astype(np.float32)
    h = (0.1, 0.2, 0.3)
    kw = dict(accuracy=4, step_size=h)
    call = (lambda a: fdx.divergence(a, keepdims=False, **kw)) if kind == "divergence" else (lambda a: getattr(fdx, kind)(a, **kw))
    before = _cabi.launch_count()
    out_host = call(x)
    launches = _cabi.launch_count() - before
    assert isinstance(out_host, np.ndarray) and launches >= 2, launches  # several slabs
    out_dev = call(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(out_host, out_dev)
    # pinned torch tensor in -> pinned torch tensor out
    xt = torch.from_numpy(x).pin_memory()
    r = call(xt)
    assert isinstance(r, torch.Tensor) and not r.is_cuda and torch.equal(r, torch.from_numpy(out_dev))


def test_host_streaming_fp64_against_the_oracle():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((272, 256, 256))
    out = fdx.laplacian(x, accuracy=2, step_size=0.5)
    sub = slice(100, 140)
    ref = O.laplacian(x[92:148], accuracy=2, step_size=0.5)[8:48]
    np.testing.assert_allclose(out[sub], ref, rtol=1e-12, atol=1e-9)
    ref0 = O.laplacian(x[:24], accuracy=2, step_size=0.5)[:12]
    np.testing.assert_allclose(out[:12], ref0, rtol=1e-12, atol=1e-9)


def test_eight_slabs_of_a_few_tile_grid_equal_the_public_operator():
    """512 x 80 x 192 fp32 laplacian in 8 slabs (tools/dist_check.py at 8 ranks, emulated): few tiles and a long z axis.
    The whole-grid call once overflowed the table of z cuts and fell back to the three-pass composition, which differs
    from the fused slab launches in the last bit; both must be the fused kernel and agree exactly."""
    torch.manual_seed(3)
    world, n1, n2, acc, h = 8, 80, 192, 4, (0.1, 0.2, 0.3)
    n0 = 64 * world
    x = torch.randn((n0, n1, n2), device="cuda")
    ref = fdx.laplacian(x, accuracy=acc, step_size=h)
    assert _cabi.last_kernel() == "fused_laplacian3d", _cabi.last_kernel()
    spec = fs.make_spec("laplacian", (n0, n1, n2), accuracy=acc, step_size=h)
    P = spec.radius
    per = n0 // world
    for r in range(world):
        z0, z1 = r * per, (r + 1) * per
        ins = torch.zeros((per + 2 * P, n1, n2), device="cuda")
        lo_b, hi_b = max(0, z0 - P), min(n0, z1 + P)
        ins[lo_b - (z0 - P) : hi_b - (z0 - P)] = x[lo_b:hi_b]
        out = torch.empty((per, n1, n2), device="cuda")
        lo, hi = z0 + (P if r > 0 else 0), z1 - (P if r < world - 1 else 0)
        for a, b in ((lo, hi), (z0, lo), (hi, z1)):
            if b > a:
                fs.run_slab(spec, [ins], z0 - P, [out], z0, a, b)
        assert torch.equal(out, ref[z0:z1]), r
