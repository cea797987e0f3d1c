"""One-GPU emulation of tools/dist_check.py: every rank's interior + shell launches against the whole-grid operator."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import slab as S

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
bad = 0
for kind, dtype, acc, n0, n1, n2 in (("laplacian", torch.float32, 4, 64 * world, 80, 192), ("divergence", torch.float64, 2, 40 * world, 48, 96),
                                     ("curl", torch.float32, 4, 48 * world, 64, 128), ("gradient", torch.float32, 6, 56 * world, 48, 200)):
    h = (0.1, 0.2, 0.3)
    nin, nout = S.FIELDS[kind]
    x = torch.randn((nin, n0, n1, n2), device="cuda", dtype=dtype)
    if kind == "laplacian":
        ref = fdx.laplacian(x[0], accuracy=acc, step_size=h)[None]
    elif kind == "divergence":
        ref = fdx.divergence(x, accuracy=acc, step_size=h, keepdims=False)[None]
    elif kind == "gradient":
        ref = fdx.gradient(x[0], accuracy=acc, step_size=h)
    else:
        ref = fdx.curl(x, accuracy=acc, step_size=h)
    spec = S.make_spec(kind, (n0, n1, n2), accuracy=acc, step_size=h)
    P = spec.radius
    per = n0 // world
    for r in range(world):
        z0, z1 = r * per, (r + 1) * per
        lo_b, hi_b = max(0, z0 - P), min(n0, z1 + P)
        ins = [torch.zeros((per + 2 * P, n1, n2), device="cuda", dtype=dtype) for _ in range(nin)]
        for f in range(nin):
            ins[f][lo_b - (z0 - P) : hi_b - (z0 - P)] = x[f, lo_b:hi_b]
        outs = [torch.empty((per, n1, n2), device="cuda", dtype=dtype) for _ in range(nout)]
        lo = z0 + (P if r > 0 else 0)
        hi = z1 - (P if r < world - 1 else 0)
        S.run_slab(spec, ins, z0 - P, outs, z0, lo, hi)
        if r > 0:
            S.run_slab(spec, ins, z0 - P, outs, z0, z0, lo)
        if r < world - 1:
            S.run_slab(spec, ins, z0 - P, outs, z0, hi, z1)
        for o in range(nout):
            d = (outs[o] - ref[o, z0:z1]).abs()
            if d.max().item() != 0:
                bad += 1
                idx = torch.nonzero(d)
                print(f"{kind} rank {r} output {o}: {idx.shape[0]} points differ, max {d.max().item():.3e}; first {idx[:4].tolist()} (local z; slab planes {z0}..{z1}, interior {lo}..{hi})")
print("slab_repro", "OK" if bad == 0 else f"{bad} mismatching outputs")

# --- where do the laplacian differences sit?  (rank 3 of 8)
if bad:
    kind, dtype, acc, n0, n1, n2 = "laplacian", torch.float32, 4, 64 * world, 80, 192
    h = (0.1, 0.2, 0.3)
    x = torch.randn((1, n0, n1, n2), device="cuda", dtype=dtype)
    spec = S.make_spec(kind, (n0, n1, n2), accuracy=acc, step_size=h)
    P = spec.radius
    for env in ({}, {"FDX_FUSED_NOPLAIN": "1"}, {"FDX_FUSED_CHUNK": "16"}, {"FDX_FUSED_CHUNK": "5"}):
        for k in ("FDX_FUSED_NOPLAIN", "FDX_FUSED_CHUNK"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ref = fdx.laplacian(x[0], accuracy=acc, step_size=h)
        os.environ.pop("FDX_FUSED_NOPLAIN", None); os.environ.pop("FDX_FUSED_CHUNK", None)
        ref0 = fdx.laplacian(x[0], accuracy=acc, step_size=h)
        d = (ref - ref0).abs()
        print("whole grid", env, "vs default: differing points", int((d > 0).sum()), "max", d.max().item())
    r = 3
    per = n0 // world
    z0, z1 = r * per, (r + 1) * per
    ins = [x[0, z0 - P : z1 + P].contiguous()]
    outs = [torch.empty((per, n1, n2), device="cuda", dtype=dtype)]
    S.run_slab(spec, ins, z0 - P, outs, z0, z0, z1)
    ref0 = fdx.laplacian(x[0], accuracy=acc, step_size=h)
    d = (outs[0] - ref0[z0:z1]).abs() > 0
    print("one launch for the whole slab: differing", int(d.sum()))
    print(" by local z:", d.sum(dim=(1, 2)).tolist()[:16], "...")
    print(" by y:", d.sum(dim=(0, 2)).tolist()[:20])
    print(" by x:", d.sum(dim=(0, 1)).tolist()[:24])
    from finitediffx_b200 import finite_diff as FD
    print("alpha slab", spec.alpha)
