"""What do the chunks of z-band planes cost?  The fused laplacian on the whole 512^3 grid against the same launch without
the band planes (output planes [nl, n0 - nr) through the slab entry point)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import slab as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dtype = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else torch.float32
acc = int(sys.argv[3]) if len(sys.argv) > 3 else 4
xs = [torch.randn((n, n, n), device="cuda", dtype=dtype) for _ in range(3)]
out = torch.empty_like(xs[0])
h = 1.0 / (n - 1)
spec = S.make_spec("laplacian", (n, n, n), accuracy=acc, step_size=h)
nl = 8  # generous: well inside the main planes for accuracy <= 4 (and the reach of accuracy 8)


def t(fn, k=20):
    for i in range(5):
        fn(xs[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(k):
        fn(xs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


for rnd in range(3):
    full = t(lambda x: S.run_slab(spec, [x], 0, [out], 0, 0, n))
    mid = t(lambda x: S.run_slab(spec, [x], 0, [out], 0, nl, n - nl))
    print(f"whole grid {full*1000:.1f} us; planes [{nl}, {n - nl}) {mid*1000:.1f} us = {mid / (n - 2 * nl) * n * 1000:.1f} us scaled to {n} planes", flush=True)
