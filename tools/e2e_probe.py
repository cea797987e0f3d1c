"""PCIe probe for the e2e path: raw pinned copy bandwidth each way / both ways, and the public-API
laplacian on host tensors for a few slab sizes.  python tools/e2e_probe.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import slab as _slab, finite_diff as _fd

n = 512
h = torch.empty((n, n, n), dtype=torch.float32, pin_memory=True).normal_()
h2 = torch.empty_like(h, pin_memory=True)
d = torch.empty((n, n, n), device="cuda")
d2 = torch.randn((n, n, n), device="cuda")
nbytes = h.numel() * 4
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
print("H2D GB/s", nbytes / t(lambda: d.copy_(h, non_blocking=True)) / 1e9)
print("D2H GB/s", nbytes / t(lambda: h2.copy_(d2, non_blocking=True)) / 1e9)
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("both ways GB/s per direction", nbytes / t(both) / 1e9)
for mb in (64, 32, 16, 8, 4):
    _slab.DEFAULT_SLAB_BYTES = mb << 20
    _fd._streamers.clear()
    el = t(lambda: fdx.laplacian(h, accuracy=4, step_size=1.0 / (n - 1)), reps=5)
    print(f"slab {mb} MB: e2e {el*1e3:.2f} ms  {n**3/el/1e9:.2f} Gpts/s")
# the bench.py loop shape: two alternating inputs, the previous result still alive
_slab.DEFAULT_SLAB_BYTES = 64 << 20
_fd._streamers.clear()
hx = [h, torch.empty_like(h, pin_memory=True).normal_()]
for i in range(2):
    r = fdx.laplacian(hx[i % 2], accuracy=4, step_size=1.0 / (n - 1))
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(6):
        t1 = time.perf_counter()
        r = fdx.laplacian(hx[i % 2], accuracy=4, step_size=1.0 / (n - 1))
        print(f"   call {i}: {(time.perf_counter()-t1)*1e3:.2f} ms", end="")
    torch.cuda.synchronize()
    print(f"\nbench-shaped loop: {(time.perf_counter()-t0)/6*1e3:.2f} ms per call")
