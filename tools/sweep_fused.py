"""Tuning sweep for the fused 3-D kernels (run on the GPU box; build with FDX_SWEEP=1 first).

    FDX_SWEEP=1 python -m finitediffx_b200._build --force
    gpurun -- python tools/sweep_fused.py [op] [n] [dtype] [accuracy]

Each configuration is checked against the composed axis-kernel path before it is timed.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx  # noqa: E402
from finitediffx_b200 import _cabi  # noqa: E402

op = sys.argv[1] if len(sys.argv) > 1 else "laplacian"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dtype = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else torch.float32
acc = int(sys.argv[4]) if len(sys.argv) > 4 else 4
cfgs = [int(c) for c in os.environ.get("SWEEP_CFGS", "0,1,2,3,4,5,6,7,8,9,10,11").split(",")]
chunks = os.environ.get("SWEEP_CHUNKS", "0,32,64,128").split(",")  # "L" (uniform) or "L:S" (graded long:short)

dev = torch.device("cuda")
fields_in = 3 if op in ("divergence", "curl") else 1
fields_out = 3 if op in ("gradient", "curl") else 1
shape = (n, n, n) if fields_in == 1 else (3, n, n, n)
xs = [torch.randn(shape, device=dev, dtype=dtype) for _ in range(3)]
h = 1.0 / (n - 1)


def call(x):
    if op == "divergence":
        return fdx.divergence(x, accuracy=acc, step_size=h, keepdims=False)
    return getattr(fdx, op)(x, accuracy=acc, step_size=h)


os.environ["FDX_FUSED_DISABLE"] = "1"
__import__("finitediffx_b200._cabi", fromlist=["reload_env"]).reload_env()  # the library snapshots the FDX_* switches
ref = call(xs[0])
torch.cuda.synchronize()
ref_kernel = _cabi.last_kernel()
os.environ["FDX_FUSED_DISABLE"] = "0"
__import__("finitediffx_b200._cabi", fromlist=["reload_env"]).reload_env()  # the library snapshots the FDX_* switches
scale = ref.abs().max().item()
bytes_pp = (fields_in + fields_out) * xs[0].element_size()
results = []
for cfg in cfgs:
    for chunk in chunks:
        os.environ["FDX_FUSED_CFG"] = str(cfg)
        __import__("finitediffx_b200._cabi", fromlist=["reload_env"]).reload_env()  # the library snapshots the FDX_* switches
        os.environ["FDX_FUSED_CHUNK"] = str(chunk).split(":")[0]
        __import__("finitediffx_b200._cabi", fromlist=["reload_env"]).reload_env()  # the library snapshots the FDX_* switches
        os.environ["FDX_FUSED_CHUNK_SHORT"] = str(chunk).split(":")[1] if ":" in str(chunk) else "0"
        __import__("finitediffx_b200._cabi", fromlist=["reload_env"]).reload_env()  # the library snapshots the FDX_* switches
        try:
            out = call(xs[0])
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"cfg {cfg} chunk {chunk}: FAILED {e}")
            continue
        kern = _cabi.last_kernel()
        err = ((out - ref).abs().max() / scale).item()
        del out
        for i in range(3):
            call(xs[i % 3])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for i in range(reps):
            call(xs[i % 3])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = bytes_pp * n**3 / ms / 1e6
        r = dict(op=op, n=n, cfg=cfg, chunk=chunk, ms=round(ms, 4), gpts=round(n**3 / ms / 1e6, 1), gbs=round(gbs, 1), err=err, kernel=kern)
        results.append(r)
        print(json.dumps(r), flush=True)
best = max(results, key=lambda r: r["gpts"]) if results else None
print("BEST", json.dumps(best), "reference kernel:", ref_kernel)
