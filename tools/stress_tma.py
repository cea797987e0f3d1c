"""Randomised stress of the persistent TMA-ring kernels (forced onto small arrays): `difference` along every axis and the
2-D operators must equal the non-persistent kernels bit for bit.

    python tools/stress_tma.py [cases] [seed]
"""
import os, sys, random
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ON = dict(FDX_ROWTMA_MIN_MB="0", FDX_COLTMA_MIN_MB="0", FDX_FUSED2D_TMA_MIN_MB="0")
OFF = dict(FDX_ROWTMA_DISABLE="1", FDX_COLTMA_DISABLE="1", FDX_FUSED2D_TMA_DISABLE="1")


def env(e):
    for k in list(ON) + list(OFF):
        os.environ.pop(k, None)
    os.environ.update(e)
    _cabi.reload_env()


bad = tma = 0
for case in range(cases):
    dtype = rnd.choice([torch.float32, torch.float64])
    V = 4 if dtype == torch.float32 else 2
    method = rnd.choice(["central", "forward", "backward"])
    d, a = rnd.randint(1, 4), rnd.randint(1, 8)
    if method != "central" and d + a > 12:
        a = 12 - d
    need = 2 * (d + a) + 2
    nd = rnd.choice([2, 2, 3, 4])
    if rnd.random() < 0.35 and nd == 2:  # a 2-D operator
        op = rnd.choice(["laplacian", "gradient", "divergence", "curl"])
        acc = rnd.choice([1, 2, 3, 4, 6, 8]) if method == "central" else rnd.choice([1, 2, 3, 4])
        dd = 2 if op == "laplacian" else 1
        if method != "central" and dd + acc - 1 > 6:
            acc = 2
        need = 2 * (dd + acc) + 2
        shape = (rnd.randint(need, 150), V * rnd.randint(32, 90))
        x = torch.randn(shape if op in ("laplacian", "gradient") else (2,) + shape, device="cuda", dtype=dtype)
        call = lambda: getattr(fdx, op)(x, accuracy=acc, step_size=(0.1, 0.3), method=method)  # noqa: E731
        label = f"{op}2d {method} a={acc} {shape}"
    else:
        axis = rnd.randrange(nd)
        shape = [rnd.randint(1, 6) for _ in range(nd)]
        shape[axis] = rnd.randint(need, need + 120)
        if axis == nd - 1:
            shape[axis] = V * max(32, (shape[axis] + V - 1) // V + rnd.randint(0, 40))
        else:
            shape[-1] = V * rnd.randint(32, 70)
        x = torch.randn(tuple(shape), device="cuda", dtype=dtype)
        call = lambda: fdx.difference(x, axis=axis, accuracy=a, derivative=d, method=method, step_size=0.37)  # noqa: E731
        label = f"difference {method} d={d} a={a} axis={axis} {tuple(shape)}"
    env(ON)
    o1 = call()
    k1 = _cabi.last_kernel()
    env(OFF)
    o2 = call()
    k2 = _cabi.last_kernel()
    ok = torch.equal(o1, o2)
    tma += k1.endswith("tma")
    if not ok:
        bad += 1
    print(f"{case:3d} {str(dtype)[6:]:8s} {label}  [{k1} / {k2}]{'' if ok else '   <-- MISMATCH ' + str((o1 - o2).abs().max().item())}", flush=True)
env({})
print(f"stress_tma: {cases} cases, {tma} through the TMA kernels, {bad} mismatches")
sys.exit(1 if bad else 0)
