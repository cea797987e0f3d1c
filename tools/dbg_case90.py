import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi
torch.manual_seed(0)
op, dtype, acc, shape = "curl", torch.float32, 4, (3, 84, 51, 68)
if len(sys.argv) > 1:
    op = sys.argv[1]
h = (0.1, 0.2, 0.3)
x = torch.randn(shape if op in ("curl", "divergence") else shape[1:], device="cuda", dtype=dtype)
kw = dict(accuracy=acc, step_size=h)
if op == "divergence":
    kw["keepdims"] = False
def run():
    a = x.clone().requires_grad_(True)
    out = getattr(fdx, op)(a, **kw)
    g = torch.randn(out.shape, device="cuda", dtype=dtype, generator=torch.Generator(device="cuda").manual_seed(5))
    (ga,) = torch.autograd.grad(out, a, g)
    return ga, _cabi.last_kernel()
os.environ["FDX_FUSED_DISABLE"] = "1"
ref, _ = run()
os.environ["FDX_FUSED_DISABLE"] = "0"
for env in ({}, {"FDX_FUSED_NOPLAIN": "1"}, {"FDX_FUSED_ZB_MERGE": "1"}, {"FDX_FUSED_NOPLAIN": "1", "FDX_FUSED_ZB_MERGE": "1"}, {"FDX_XBAND_SLOW": "1"}, {"FDX_FUSED_RIMFIRST": "0"},
            {"FDX_FUSED_CHUNK": "16"}, {"FDX_FUSED_CHUNK": "64"}):
    for k in ("FDX_FUSED_NOPLAIN", "FDX_FUSED_ZB_MERGE", "FDX_XBAND_SLOW", "FDX_FUSED_RIMFIRST", "FDX_FUSED_CHUNK"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ga, k = run()
    d = (ga - ref).abs()
    idx = torch.nonzero(d > 1e-3 * ref.abs().max())
    print(env, k, "max err", (d.max() / ref.abs().max()).item(), "bad points", idx.shape[0], "first", idx[:3].tolist(), "last", idx[-3:].tolist())
