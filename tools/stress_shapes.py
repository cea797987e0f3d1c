"""Randomised shape stress: fused kernels (all marches) against the composition of axis kernels, forward and vjp.

    python tools/stress_shapes.py [cases] [seed]
"""
import os, sys, random
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
fusedn = 0
for case in range(cases):
    dtype = rnd.choice([torch.float32, torch.float64])
    V = 4 if dtype == torch.float32 else 2
    acc = rnd.choice([2, 2, 4, 4, 6, 8])
    lo = 2 * acc + 2  # the adjoint bands need p + L rows per side
    mode = rnd.random()
    if mode < 0.25:  # thin in one or two directions
        n0, n1, n2 = rnd.randint(lo, lo + 6), rnd.randint(lo, 130), V * rnd.randint((lo + V - 1) // V, 330 // V)
    elif mode < 0.5:  # around the tile sizes (64 x 16 fp32, 32 x 16 fp64) and their multiples
        n0 = rnd.randint(lo, 70)
        n1 = rnd.choice([15, 16, 17, 31, 32, 33, 47, 48, 49, 63, 64, 65]) + rnd.choice([0, 0, 16])
        n2 = V * rnd.choice([8, 15, 16, 17, 18, 31, 32, 33, 34, 48, 49, 64, 65]) * rnd.choice([1, 1, 2])
        n1, n2 = max(n1, lo), max(n2, V * ((lo + V - 1) // V))
    else:
        n0 = rnd.randint(lo, 90)
        n1 = rnd.randint(lo, 130)
        n2 = V * rnd.randint((lo + V - 1) // V, 330 // V)
    op = rnd.choice(["laplacian", "gradient", "divergence", "curl"])
    h = (rnd.uniform(0.05, 0.5), rnd.uniform(0.05, 0.5), rnd.uniform(0.05, 0.5))
    shape = (n0, n1, n2) if op in ("laplacian", "gradient") else (3, n0, n1, n2)
    x = torch.randn(shape, device="cuda", dtype=dtype)
    kw = dict(accuracy=acc, step_size=h)
    if op == "divergence":
        kw["keepdims"] = False

    def run():
        a = x.clone().requires_grad_(True)
        out = getattr(fdx, op)(a, **kw)
        k = _cabi.last_kernel()
        g = torch.randn(out.shape, device="cuda", dtype=dtype, generator=torch.Generator(device="cuda").manual_seed(case))
        (ga,) = torch.autograd.grad(out, a, g)
        return out.detach(), ga, k, _cabi.last_kernel()

    os.environ["FDX_FUSED_DISABLE"] = "0"
    __import__("finitediffx_b200._cabi", fromlist=["reload_env"]).reload_env()  # the library snapshots the FDX_* switches
    out, ga, k1, k2 = run()
    os.environ["FDX_FUSED_DISABLE"] = "1"
    __import__("finitediffx_b200._cabi", fromlist=["reload_env"]).reload_env()  # the library snapshots the FDX_* switches
    out_r, ga_r, _, _ = run()
    os.environ["FDX_FUSED_DISABLE"] = "0"
    __import__("finitediffx_b200._cabi", fromlist=["reload_env"]).reload_env()  # the library snapshots the FDX_* switches
    tol = 3e-5 if dtype == torch.float32 else 2e-11
    e1 = ((out - out_r).abs().max() / out_r.abs().max()).item()
    e2 = ((ga - ga_r).abs().max() / ga_r.abs().max()).item()
    fusedn += k1.startswith("fused_") + k2.startswith("fused_")
    flag = "" if (e1 <= tol and e2 <= tol) else "   <-- MISMATCH"
    if flag:
        bad += 1
    print(f"{case:3d} {op:10s} {str(dtype)[6:]:8s} a={acc} shape={n0}x{n1}x{n2}  fwd {e1:.2e} [{k1}]  vjp {e2:.2e} [{k2}]{flag}", flush=True)
torch.cuda.synchronize()
print(f"stress: {cases} cases, {fusedn} fused launches checked, {bad} mismatches")
sys.exit(1 if bad else 0)
