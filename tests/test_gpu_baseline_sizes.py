"""GPU parity at the sizes BASELINE.json names (configs 2-5), against the oracle.

The oracle cannot run a 512^3 / 1024^3 / 2048^2-plane grid in seconds, so each case compares
*sub-bricks* (corners, faces, edges and the centre: every combination of one-sided and central
stencils) cut out of the full-size CUDA result with the oracle run on the same brick of the input plus
a margin: bricks that touch a physical boundary keep it, artificial edges are cut away.  The vjp
cases compare with the dense transposed matrices of the oracle (`jax.grad` through the reference is
exactly D^T) applied to pencils of the cotangent through the brick.

Tolerances (north star): fp32 <= 1e-5, fp64 <= 1e-12 on the conditioning scale (tests/_parity.py).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi
from finitediffx_b200 import slab as fs
from oracle import fdx_oracle as O
from tests._parity import TOL, scale_for

pytestmark = pytest.mark.gpu

PAD = 10  # margin around a brick: >= the longest one-sided stencil reach of the configs (d=2, a=8: 9)

# out[co] += sign * D(axis, derivative) in[ci]
TERMS = {
    "laplacian": [(0, 0, k, +1, 2) for k in range(3)],
    "gradient": [(k, 0, k, +1, 1) for k in range(3)],
    "divergence": [(0, k, k, +1, 1) for k in range(3)],
    "curl": [(0, 2, 1, +1, 1), (0, 1, 2, -1, 1), (1, 0, 2, +1, 1), (1, 2, 0, -1, 1), (2, 1, 0, +1, 1), (2, 0, 1, -1, 1)],
}
N_IN = {"laplacian": 1, "gradient": 1, "divergence": 3, "curl": 3}
N_OUT = {"laplacian": 1, "gradient": 3, "divergence": 1, "curl": 3}


def _bricks(shape, size=20):
    """Corner, edge, face and centre bricks of a 3-D grid (as tuples of slices)."""
    def pos(n):
        return (slice(0, size), slice(n // 2 - size // 2, n // 2 - size // 2 + size), slice(n - size, n))

    pz, py, px = pos(shape[0]), pos(shape[1]), pos(shape[2])
    picks = [(0, 0, 0), (2, 2, 2), (1, 1, 1), (0, 1, 2), (2, 0, 1), (1, 2, 0), (0, 2, 0), (1, 1, 2), (2, 1, 1)]
    return [(pz[a], py[b], px[c]) for a, b, c in picks]


def _ext(brick, shape):
    return tuple(slice(max(s.start - PAD, 0), min(s.stop + PAD, n)) for s, n in zip(brick, shape))


def _fields(t, nf):
    """(nf, n0, n1, n2) view of a scalar (n0, n1, n2) or vector field tensor."""
    return t[None] if t.ndim == 3 else t[:nf]


def check_forward_bricks(op, x, out, kw, tol, bricks=None):
    """x, out: full-size CUDA tensors (the op's public input / output)."""
    shape = tuple(x.shape[-3:])
    worst = 0.0
    for brick in bricks or _bricks(shape):
        ext = _ext(brick, shape)
        core = tuple(slice(b.start - e.start, b.stop - e.start) for b, e in zip(brick, ext))
        xb = (x[ext] if x.ndim == 3 else x[(slice(None),) + ext]).cpu().numpy()
        ref = getattr(O, op)(xb, **kw)
        sc = scale_for(op, xb, kw)
        lead = (slice(None),) * (ref.ndim - 3)
        got = (out[brick] if out.ndim == 3 else out[(slice(None),) + brick]).cpu().numpy()
        r, s = ref[lead + core], np.broadcast_to(sc, ref.shape)[lead + core]
        assert got.shape == r.shape, (op, got.shape, r.shape)
        err = float(np.max(np.abs(got.astype(np.float64) - r.astype(np.float64)) / np.where(s == 0, 1.0, s)))
        worst = max(worst, err)
        assert err <= tol, (op, kw, brick, err)
    return worst


def check_vjp_bricks(op, g, gx, kw, tol, bricks=None):
    """gx = A^T g (CUDA tensors, full size) against the oracle's dense transposed matrices."""
    shape = tuple(g.shape[-3:])
    acc, h, method = kw["accuracy"], kw["step_size"], kw.get("method", "central")
    hs = h if isinstance(h, (tuple, list)) else (h,) * 3
    gf, xf = _fields(g, N_OUT[op]), _fields(gx, N_IN[op])
    dense = {}
    worst = 0.0
    for brick in bricks or _bricks(shape, size=12):
        want = np.zeros((N_IN[op],) + tuple(s.stop - s.start for s in brick))
        scale = np.zeros_like(want)
        for co, ci, ax, sign, d in TERMS[op]:
            key = (ax, d)
            if key not in dense:
                dense[key] = O.dense_matrix(shape[ax], accuracy=acc, derivative=d, method=method, step_size=hs[ax])
            rows = dense[key].T[brick[ax]]  # rows of D^T that produce the brick's range along `ax`
            idx = list(brick)
            idx[ax] = slice(None)
            pencil = gf[co][tuple(idx)].double().cpu().numpy()
            want[ci] += sign * np.moveaxis(np.tensordot(rows, np.moveaxis(pencil, ax, 0), axes=1), 0, ax)
            scale[ci] += np.moveaxis(np.tensordot(np.abs(rows), np.abs(np.moveaxis(pencil, ax, 0)), axes=1), 0, ax)
        got = torch.stack([xf[ci][brick] for ci in range(N_IN[op])]).double().cpu().numpy()
        err = float(np.max(np.abs(got - want) / np.where(scale == 0, 1.0, scale)))
        worst = max(worst, err)
        assert err <= tol, (op, kw, brick, err)
    return worst


def _call(op, x, kw):
    if op == "divergence":
        return fdx.divergence(x, keepdims=False, **kw)
    return getattr(fdx, op)(x, **kw)


def _okw(op, kw):
    return dict(kw, keepdims=False) if op == "divergence" else kw


# ------------------------------------------------ operator-level vjp against the oracle (small grids, every accuracy)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_operator_level_vjp_equals_the_oracles_transposed_matrices(dtype):
    """vjp of every fused 3-D operator = sum of D_k^T g over its terms with D_k the oracle's dense matrix (what jax.grad
    through the reference computes), on the whole grid, for every radius the fused kernels cover and all three methods."""
    shape = (30, 34, 40)
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    whole = [tuple(slice(0, n) for n in shape)]
    gen = torch.Generator(device="cuda").manual_seed(9)
    for method, accs in (("central", (1, 2, 4, 6, 8, 10)), ("forward", (1, 3, 5)), ("backward", (2, 4))):
        for acc in accs:
            for op in ("laplacian", "gradient", "divergence", "curl"):
                d = 2 if op == "laplacian" else 1
                if method != "central" and d + acc - 1 > 6:
                    continue
                kw = dict(accuracy=acc, step_size=(0.1, 0.2, 0.3), method=method)
                x = torch.randn(shape if N_IN[op] == 1 else (3,) + shape, device="cuda", dtype=dtype, generator=gen).requires_grad_(True)
                out = _call(op, x, kw)
                g = torch.randn(out.shape, device="cuda", dtype=dtype, generator=gen)
                (gx,) = torch.autograd.grad(out, x, g)
                assert _cabi.last_kernel().startswith("fused_"), (op, kw, _cabi.last_kernel())
                check_vjp_bricks(op, g, gx, kw, tol, bricks=whole)


# ------------------------------------------------------------------------- config 3: 512^3 fp32
@pytest.mark.parametrize("op", ["laplacian", "gradient", "divergence", "curl"])
def test_config3_512cubed_fp32_forward_and_vjp(op):
    n = 512
    kw = dict(accuracy=4, step_size=1.0 / (n - 1))
    gen = torch.Generator(device="cuda").manual_seed(0)
    shape = (n, n, n) if N_IN[op] == 1 else (3, n, n, n)
    x = torch.randn(shape, device="cuda", generator=gen).requires_grad_(True)
    out = _call(op, x, kw)
    assert _cabi.last_kernel().startswith("fused_"), _cabi.last_kernel()
    w = check_forward_bricks(op, x.detach(), out.detach(), _okw(op, kw), 1e-5)
    g = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    (gx,) = torch.autograd.grad(out, x, g)
    assert _cabi.last_kernel().startswith("fused_"), _cabi.last_kernel()
    wv = check_vjp_bricks(op, g, gx, kw, 1e-5)
    print(f"config 3 {op}: forward {w:.2e}, vjp {wv:.2e}")


# ---------------------------------------------------------------- config 4: (3, 1024^3) fp32
@pytest.mark.parametrize("op", ["divergence", "curl"])
def test_config4_1024cubed_fp32(op):
    n = 1024
    kw = dict(accuracy=4, step_size=1.0 / (n - 1))
    x = torch.empty((3, n, n, n), device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(2)
    for c in range(3):
        x[c].normal_(generator=gen)
    out = _call(op, x, kw)
    assert _cabi.last_kernel().startswith("fused_"), _cabi.last_kernel()
    w = check_forward_bricks(op, x, out, _okw(op, kw), 1e-5)
    print(f"config 4 {op}: {w:.2e}")


# ------------------------------------------------- config 5: fp64 accuracy 8 on 2048^2 planes
def test_config5_fp64_accuracy8_slab_of_2048cubed():
    """One rank's share of the 2048^3 fp64 grid at 8 GPUs (256 planes + 5-plane halos = 266), twice: as a grid of
    its own (one-sided stencils at both z ends) and as slab 3 of 8 of the global grid through the slab entry point
    (global plane indices, halos in the buffer, central stencils at the cuts)."""
    n, acc = 2048, 8
    h = 1.0 / (n - 1)
    kw = dict(accuracy=acc, step_size=h)
    x = torch.empty((266, n, n), device="cuda", dtype=torch.float64)
    x.normal_(generator=torch.Generator(device="cuda").manual_seed(3))
    out = fdx.laplacian(x, **kw)
    assert _cabi.last_kernel() == "fused_laplacian3d", _cabi.last_kernel()
    w = check_forward_bricks("laplacian", x, out, kw, 1e-12)
    # the same buffer as planes [763, 1029) of the global grid: outputs [768, 1024)
    spec = fs.make_spec("laplacian", (n, n, n), accuracy=acc, step_size=h)
    P = spec.radius
    assert P == 5
    z0, z1 = 768, 1024
    res = torch.empty((z1 - z0, n, n), device="cuda", dtype=torch.float64)
    fs.run_slab(spec, [x], z0 - P, [res], z0, z0, z1)
    # oracle bricks live in buffer coordinates; they must stay clear of the buffer's ends (artificial edges)
    bricks = [b for b in _bricks((266, n, n)) if b[0].start >= PAD + P and b[0].stop <= 266 - PAD - P]
    shifted = torch.empty_like(x)  # slab outputs at their buffer positions
    shifted[P : P + (z1 - z0)] = res
    w2 = check_forward_bricks("laplacian", x, shifted, kw, 1e-12, bricks=bricks)
    assert len(bricks) >= 3
    print(f"config 5: grid {w:.2e}, slab of the global grid {w2:.2e}")


# ------------------------------------------------------------- config 2: 8192^2 fp32, 168 cases
def test_config2_all_168_cases_8192_squared():
    """derivative 1-4 x accuracy 2-8 x forward/backward/central x both axes on (8192, 8192) fp32.  The 1-D operator
    acts on rows (axis 1) or columns (axis 0) independently: the oracle runs on 32 strided rows / columns."""
    n = 8192
    x = torch.randn((n, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    pick = torch.arange(5, n, 257, device="cuda")
    sub = {0: x[:, pick].cpu().numpy(), 1: x[pick, :].cpu().numpy()}
    worst, cases, kernels = 0.0, 0, set()
    for method in O.METHODS:
        for d in range(1, 5):
            for a in range(2, 9):
                for axis in (0, 1):
                    kw = dict(axis=axis, accuracy=a, derivative=d, method=method, step_size=0.5)
                    out = fdx.difference(x, **kw)
                    kernels.add(_cabi.last_kernel())
                    got = (out[:, pick] if axis == 0 else out[pick, :]).cpu().numpy()
                    ref = O.difference(sub[axis], **kw)
                    sc = scale_for("difference", sub[axis], kw)
                    err = float(np.max(np.abs(got.astype(np.float64) - ref) / np.where(sc == 0, 1.0, sc)))
                    worst = max(worst, err)
                    assert err <= 1e-5, (method, d, a, axis, err, _cabi.last_kernel())
                    cases += 1
    assert cases == 168
    assert kernels <= {"axis_stream", "axis_row", "axis_rowtma", "axis_coltma"}, kernels  # the vector kernels, never a scalar fallback
    print(f"config 2: 168 cases, worst parity error {worst:.2e}")


# ------------------------------------------------------------------ multi-GPU: real ranks, NCCL
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (the single-GPU emulation is tests/test_gpu_slab.py)")
def test_slab_operator_on_real_ranks_equals_the_single_gpu_operator():
    """All four fused operators, slab-decomposed over every GPU of the box with NCCL halo exchange, must be
    bit-identical to the single-GPU operator on the gathered grid (tools/dist_check.py under torchrun)."""
    world = min(8, torch.cuda.device_count())
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(root, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0 and "dist_check OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
