"""Run one operator a few times (for ncu): python tools/prof_one.py [op] [n] [dtype] [accuracy] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx  # noqa: E402

op = sys.argv[1] if len(sys.argv) > 1 else "laplacian"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dtype = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else torch.float32
acc = int(sys.argv[4]) if len(sys.argv) > 4 else 4
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
fields_in = 3 if op.split("_")[0] in ("divergence", "curl") else 1
shape = (n, n, n) if fields_in == 1 else (3, n, n, n)
xs = [torch.randn(shape, device="cuda", dtype=dtype) for _ in range(2)]
h = 1.0 / (n - 1)
if op.endswith("_vjp"):
    base = op[:-4]
    x = xs[0].requires_grad_(True)
    out = fdx.divergence(x, accuracy=acc, step_size=h, keepdims=False) if base == "divergence" else getattr(fdx, base)(x, accuracy=acc, step_size=h)
    g = torch.randn_like(out)
    for i in range(reps):
        (gx,) = torch.autograd.grad(out, x, g, retain_graph=True)
    torch.cuda.synchronize()
    print("ok", op, n, dtype, float(gx.flatten()[0]))
    sys.exit(0)
for i in range(reps):
    if op == "divergence":
        out = fdx.divergence(xs[i % 2], accuracy=acc, step_size=h, keepdims=False)
    else:
        out = getattr(fdx, op)(xs[i % 2], accuracy=acc, step_size=h)
torch.cuda.synchronize()
print("ok", op, n, dtype, float(out.flatten()[0]))
