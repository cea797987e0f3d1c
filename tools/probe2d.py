"""2-D fused operators on (8192, 8192) fp32: fused launch against the composition of axis kernels."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi
x = torch.randn(8192, 8192, device="cuda")
f = torch.randn(2, 8192, 8192, device="cuda")
def t(call):
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
for name, call, fields in (("laplacian a4", lambda: fdx.laplacian(x, accuracy=4, step_size=0.1), 2), ("laplacian a8", lambda: fdx.laplacian(x, accuracy=8, step_size=0.1), 2),
                           ("gradient a4", lambda: fdx.gradient(x, accuracy=4, step_size=0.1), 3), ("divergence a4", lambda: fdx.divergence(f, accuracy=4, step_size=0.1), 3),
                           ("curl a2", lambda: fdx.curl(f, accuracy=2, step_size=0.1), 3), ("laplacian fwd a3", lambda: fdx.laplacian(x, accuracy=3, step_size=0.1, method="forward"), 2)):
    row = {}
    for mode in ("0", "1"):
        os.environ["FDX_FUSED2D_DISABLE"] = mode; _cabi.reload_env()
        ms = t(call)
        row["composed" if mode == "1" else "fused"] = dict(ms=round(ms, 4), gbs=round(8192 * 8192 * 4 * fields / ms / 1e6), kernel=_cabi.last_kernel())
    print(name, json.dumps(row), flush=True)
