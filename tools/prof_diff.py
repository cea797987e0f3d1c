"""Run `difference` on (8192, 8192) a few times (for ncu): python tools/prof_diff.py method derivative accuracy axis [dtype] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx  # noqa: E402

method, d, a, axis = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dtype = torch.float64 if (len(sys.argv) > 5 and sys.argv[5] == "f64") else torch.float32
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 4
n = int(os.environ.get("N", 8192))
xs = [torch.randn(n, n, device="cuda", dtype=dtype) for _ in range(2)]
for i in range(reps):
    out = fdx.difference(xs[i % 2], axis=axis, accuracy=a, derivative=d, method=method, step_size=0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    out = fdx.difference(xs[i % 2], axis=axis, accuracy=a, derivative=d, method=method, step_size=0.5)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print("ok", method, d, a, axis, dtype, f"{t:.4f} ms  {n * n * 2 * xs[0].element_size() / t / 1e6:.0f} GB/s")
