"""Throughput of a fused operator on a list of grid shapes (power-of-two plane strides against others).

    python tools/shape_probe.py [op] [acc] [f32|f64] n0xn1xn2 ...
Shapes run interleaved for several rounds (clock drift hits all alike); the median is reported.
"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx  # noqa: E402
from finitediffx_b200 import _cabi  # noqa: E402

args = sys.argv[1:]
op = args[0] if args else "laplacian"
acc = int(args[1]) if len(args) > 1 else 4
dt = torch.float64 if (len(args) > 2 and args[2] == "f64") else torch.float32
shapes = [tuple(int(v) for v in s.split("x")) for s in args[3:]] or [(512, 512, 512)]
fin = 3 if op in ("divergence", "curl") else 1
fout = 3 if op in ("gradient", "curl") else 1
rounds = int(os.environ.get("ROUNDS", "5"))
reps = int(os.environ.get("REPS", "20"))


def call(x):
    if op == "divergence":
        return fdx.divergence(x, accuracy=acc, step_size=0.01, keepdims=False)
    return getattr(fdx, op)(x, accuracy=acc, step_size=0.01)


times = {s: [] for s in shapes}
kern = {}
for rnd in range(rounds):
    for s in shapes:
        xs = [torch.randn(((fin,) if fin > 1 else ()) + s, device="cuda", dtype=dt) for _ in range(3)]
        for i in range(3):
            call(xs[i])
        kern[s] = _cabi.last_kernel()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            call(xs[i % 3])
        e1.record()
        torch.cuda.synchronize()
        times[s].append(e0.elapsed_time(e1) / reps)
        del xs
for s in shapes:
    t = statistics.median(times[s])
    pts = s[0] * s[1] * s[2]
    gbs = pts * (fin + fout) * (4 if dt == torch.float32 else 8) / t / 1e6
    print(json.dumps(dict(op=op, acc=acc, dtype=str(dt)[6:], shape=list(s), ms_med=round(t, 4), ms_min=round(min(times[s]), 4),
                          gpts=round(pts / t / 1e6, 1), gbs=round(gbs, 1), kernel=kern[s])), flush=True)
