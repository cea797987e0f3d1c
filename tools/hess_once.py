"""A few calls of the 3-D hessian (for ncu captures): python tools/hess_once.py [n] [accuracy] [f32|f64]"""
import sys

import torch

import finitediffx_b200 as fdx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
acc = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dt = torch.float64 if len(sys.argv) > 3 and sys.argv[3] == "f64" else torch.float32
x = torch.randn((n, n, n), device="cuda", dtype=dt)
for _ in range(3):
    h = fdx.hessian(x, accuracy=acc, step_size=0.1)
torch.cuda.synchronize()
print(float(h[1, 1, n // 2, n // 2, n // 2]))
