"""Specialised (interior-CTA) march vs the generic march of the fused kernels: bit-equality and timing.

    python tools/plain_check.py [n]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cases = [("laplacian", torch.float32, 4), ("divergence", torch.float32, 4), ("gradient", torch.float32, 4), ("curl", torch.float32, 4),
         ("laplacian", torch.float64, 8), ("laplacian", torch.float32, 2), ("laplacian", torch.float32, 6), ("divergence", torch.float32, 6),
         ("laplacian", torch.float64, 4), ("divergence", torch.float64, 4)]
if len(sys.argv) > 2:
    cases = [c for c in cases if c[0] in sys.argv[2:]]


def call(op, x, acc, h):
    if op == "divergence":
        return fdx.divergence(x, accuracy=acc, step_size=h, keepdims=False)
    return getattr(fdx, op)(x, accuracy=acc, step_size=h)


for op, dt, acc in cases:
    fin = 3 if op in ("divergence", "curl") else 1
    fout = 3 if op in ("gradient", "curl") else 1
    shape = (n, n, n) if fin == 1 else (3, n, n, n)
    xs = [torch.randn(shape, device="cuda", dtype=dt) for _ in range(3)]
    h = 1.0 / (n - 1)
    res = {}
    outs = {}
    for mode in ("1", "0"):
        os.environ["FDX_FUSED_NOPLAIN"] = mode
        outs[mode] = call(op, xs[0], acc, h)
        for i in range(3):
            call(op, xs[i % 3], acc, h)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for i in range(reps):
            call(op, xs[i % 3], acc, h)
        e1.record()
        torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / reps
    equal = bool(torch.equal(outs["0"], outs["1"]))
    b = (fin + fout) * xs[0].element_size() * n**3
    print(json.dumps(dict(op=op, dtype=str(dt), acc=acc, n=n, bit_equal=equal, ms_generic=round(res["1"], 4), ms_plain=round(res["0"], 4),
                          gpts_plain=round(n**3 / res["0"] / 1e6, 1), gbs_plain=round(b / res["0"] / 1e6, 1), speedup=round(res["1"] / res["0"], 3))), flush=True)
    del xs, outs
