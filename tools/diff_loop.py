"""A/B of `difference` variants (env switches) on (8192, 8192) under a sustained loop, with SM clocks.

    python tools/diff_loop.py method derivative accuracy axis [reps] -- "name:K=V,..." ...
"""
import os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi
import pynvml
pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(0)
args = sys.argv[1:]
specs = []
if "--" in args:
    i = args.index("--"); specs = args[i + 1:]; args = args[:i]
method, d, a, axis = args[0], int(args[1]), int(args[2]), int(args[3])
reps = int(args[4]) if len(args) > 4 else 200
variants = [("default", {})]
for sp in specs:
    name, _, kv = sp.partition(":")
    variants.append((name, dict(item.split("=") for item in kv.split(",") if item)))
keys = set(k for _, e in variants for k in e)
n = 8192
xs = [torch.randn(n, n, device="cuda") for _ in range(2)]
out = torch.empty_like(xs[0])
res = {name: [] for name, _ in variants}
clk = {name: [] for name, _ in variants}
for rnd in range(5):
    for name, env in variants:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        _cabi.reload_env()
        for i in range(5):
            fdx.difference(xs[i % 2], axis=axis, accuracy=a, derivative=d, method=method, step_size=0.5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fdx.difference(xs[i % 2], axis=axis, accuracy=a, derivative=d, method=method, step_size=0.5)
        e1.record()
        c = pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
        torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / reps)
        clk[name].append(c)
for name, _ in variants:
    t = statistics.median(res[name])
    print(name, f"median {t*1000:.1f} us  {n*n*8/t/1e6:.0f} GB/s  min {min(res[name])*1000:.1f} us  clocks {clk[name]}  kernel {_cabi.last_kernel()}")
