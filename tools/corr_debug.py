"""Debug: vjp with the bands in correction form vs dense bands (FDX_FUSED_NOCORR=1): where do they differ?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi

op = sys.argv[1] if len(sys.argv) > 1 else "divergence"
acc = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dtype = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else torch.float32
shape = tuple(int(v) for v in sys.argv[4].split(",")) if len(sys.argv) > 4 else (30, 34, 40)
nin = 3 if op in ("divergence", "curl") else 1
gen = torch.Generator(device="cuda").manual_seed(9)
x = torch.randn(shape if nin == 1 else (3,) + shape, device="cuda", dtype=dtype, generator=gen).requires_grad_(True)
kw = dict(accuracy=acc, step_size=(0.1, 0.2, 0.3))
if op == "divergence":
    kw["keepdims"] = False
out = getattr(fdx, op)(x, **kw)
g = torch.randn(out.shape, device="cuda", dtype=dtype, generator=gen)
res = {}
for name, env in (("corr", "0"), ("dense", "1")):
    os.environ["FDX_FUSED_NOCORR"] = env
    _cabi.reload_env()
    (gx,) = torch.autograd.grad(out, x, g, retain_graph=True)
    res[name] = gx.clone()
    print(name, _cabi.last_kernel())
d = (res["corr"] - res["dense"]).abs()
sc = res["dense"].abs().max().item()
print("max diff", d.max().item(), "rel", d.max().item() / sc)
comp = d.reshape((-1,) + shape)
for k in range(comp.shape[0]):
    m = comp[k]
    idx = torch.nonzero(m > 1e-4 * sc if dtype == torch.float32 else m > 1e-11 * sc)
    print("component", k, "max", m.max().item(), "n bad", idx.shape[0])
    if idx.shape[0]:
        for ax in range(3):
            print("   axis", ax, "bad indices:", sorted(set(idx[:, ax].tolist()))[:40])
