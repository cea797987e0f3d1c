"""Does the distance between the input and the output buffer matter (DRAM bank aliasing)?

    python tools/alias_probe.py [n]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finitediffx_b200 import _cabi  # noqa: E402
from finitediffx_b200.plan import axis_plan  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nbytes = n**3 * 4
pool = torch.empty(2 * nbytes + (256 << 20), dtype=torch.uint8, device="cuda")
base = pool.data_ptr()
assert base % 256 == 0
x = pool[:nbytes].view(torch.float32).view(n, n, n)
x.normal_()
h = 1.0 / (n - 1)
plans = [axis_plan(n, 2, 4, "central", k) for k in range(3)]
alpha = [1.0 / h**2] * 3


def run(out, reps=20):
    for _ in range(3):
        _cabi.fused3d("laplacian", plans, x, out, alpha)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _cabi.fused3d("laplacian", plans, x, out, alpha)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for chunk in ("0", "256", "128"):
    os.environ["FDX_FUSED_CHUNK"] = chunk
    for delta in (0, 256, 4096, 65536, 1 << 20, (1 << 20) + 4096, 3 << 20, (33 << 20) + 8192, 100 << 20):
        out = pool[nbytes + delta : 2 * nbytes + delta].view(torch.float32).view(n, n, n)
        ms = run(out)
        print(json.dumps(dict(chunk=chunk, delta=delta, ms=round(ms, 4), gpts=round(n**3 / ms / 1e6, 1))), flush=True)
