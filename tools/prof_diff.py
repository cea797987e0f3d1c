"""A few fdx.difference calls on 8192^2 fp32 (for an ncu launch list)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
x = torch.randn(8192, 8192, device="cuda")
y = torch.empty_like(x)
y.copy_(x)
for axis in (0, 1):
    for method, d, a in (("central", 1, 2), ("central", 2, 4), ("central", 4, 8), ("forward", 1, 4), ("backward", 3, 6)):
        for _ in range(3):
            out = fdx.difference(x, axis=axis, accuracy=a, derivative=d, method=method, step_size=0.5)
torch.cuda.synchronize()
print("ok")
