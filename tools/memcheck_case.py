import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import finitediffx_b200 as fdx
torch.manual_seed(0)
for shape in [(40, 36, 72), (33, 20, 132), (16, 16, 64), (70, 17, 12), (40, 56, 200), (21, 97, 196)]:  # the last two have interior, rim and corner tiles
    for dt in (torch.float32, torch.float64):
        x = torch.randn(shape, device="cuda", dtype=dt)
        F = torch.randn((3,) + shape, device="cuda", dtype=dt)
        for acc in (2, 4, 6):
            a = fdx.laplacian(x, accuracy=acc, step_size=0.1)
            b = fdx.gradient(x, accuracy=acc, step_size=0.1)
            c = fdx.divergence(F, accuracy=acc, step_size=0.1, keepdims=False)
            d = fdx.curl(F, accuracy=acc, step_size=0.1)
        xg = x.clone().requires_grad_(True)
        fdx.laplacian(xg, accuracy=4, step_size=0.1).sum().backward()
torch.cuda.synchronize()
print("memcheck case done", fdx._cabi.last_kernel() if hasattr(fdx, "_cabi") else "")
