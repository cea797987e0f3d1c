"""Try explicit z-chunk lists (FDX_FUSED_CUTS = sizes from the END of z backwards): python tools/cuts_probe.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
n = 512; xs = [torch.randn(n, n, n, device="cuda") for _ in range(3)]; h = 1.0 / (n - 1)
ref = fdx.laplacian(xs[0], accuracy=4, step_size=h)
def t():
    for i in range(3): fdx.laplacian(xs[i % 3], accuracy=4, step_size=h)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): fdx.laplacian(xs[i % 3], accuracy=4, step_size=h)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 30
print("default", round(t(), 4))
for cuts in sys.argv[1:]:
    os.environ["FDX_FUSED_CUTS"] = cuts
    out = fdx.laplacian(xs[0], accuracy=4, step_size=h)
    ok = torch.equal(out, ref)
    print(cuts, round(t(), 4), "bit-equal" if ok else "DIFFERS")
