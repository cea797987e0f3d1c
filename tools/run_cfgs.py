"""Run the fused laplacian once per (cfg, chunk) pair given on the command line: cfg:chunk ..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx  # noqa: E402

n = 512
x = torch.randn(n, n, n, device="cuda")
h = 1.0 / (n - 1)
for pair in sys.argv[1:]:
    cfg, chunk = pair.split(":")
    os.environ["FDX_FUSED_CFG"], os.environ["FDX_FUSED_CHUNK"] = cfg, chunk
    for _ in range(2):
        out = fdx.laplacian(x, accuracy=4, step_size=h)
    torch.cuda.synchronize()
print("ok")
