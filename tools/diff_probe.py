"""BASELINE config 2 in small: `difference` on (8192, 8192) fp32 for a few stencils and both axes, under env switches.

    python tools/diff_probe.py "name:K=V,K=V" ...
"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx  # noqa: E402
from finitediffx_b200 import _cabi  # noqa: E402

variants = [("default", {})]
for sp in sys.argv[1:]:
    name, _, kv = sp.partition(":")
    variants.append((name, dict(item.split("=") for item in kv.split(",") if item)))
keys = set(k for _, e in variants for k in e)
x = torch.randn(8192, 8192, device="cuda")
cases = [("central", 1, 2), ("central", 2, 4), ("central", 1, 6), ("central", 4, 8), ("forward", 1, 4), ("backward", 3, 8)]
if os.environ.get("ALL_CASES"):  # the whole BASELINE config-2 sweep: only the medians are printed
    cases = [(m, d, a) for m in ("central", "forward", "backward") for d in range(1, 5) for a in range(2, 9)]
res = {}
for rnd in range(3):
    for name, env in variants:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        _cabi.reload_env()
        for method, d, a in cases:
            for axis in (0, 1):
                call = lambda: fdx.difference(x, axis=axis, accuracy=a, derivative=d, method=method, step_size=0.5)
                for _ in range(3):
                    call()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    call()
                e1.record()
                torch.cuda.synchronize()
                res.setdefault((name, method, d, a, axis), []).append(e0.elapsed_time(e1) / 10)
if os.environ.get("ALL_CASES"):
    for name, _ in variants:
        for axis in (0, 1):
            v = sorted(statistics.median(res[(name, m, d, a, axis)]) for m, d, a in cases)
            gb = lambda t: round(8192 * 8192 * 8 / t / 1e6)  # noqa: E731
            print(name, f"axis {axis}: median {gb(v[len(v) // 2])} GB/s, best {gb(v[0])}, worst {gb(v[-1])}", flush=True)
    sys.exit(0)
for name, _ in variants:
    row = {}
    for method, d, a in cases:
        for axis in (0, 1):
            t = statistics.median(res[(name, method, d, a, axis)])
            row[f"{method[0]}{d}{a}ax{axis}"] = round(8192 * 8192 * 8 / t / 1e6)
    print(name, json.dumps(row), flush=True)
