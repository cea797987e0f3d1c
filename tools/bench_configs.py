"""Time every BASELINE.json config on ONE B200 (device-resident, CUDA events) and print JSON lines.

    python tools/bench_configs.py [1 2 3 4 5]

Roofline fractions are against MEASURED_PEAKS.json hbm_gbs (copy) with the algorithmic bytes of
SURVEY.md section 8(d); halo / boundary re-reads are not counted.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx  # noqa: E402
from finitediffx_b200 import _cabi  # noqa: E402

PEAK = 6533.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
which = set(sys.argv[1:]) or {"1", "2", "3", "4", "5"}
dev = "cuda"


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(config, name, pts, bytes_pp, ms, **extra):
    gbs = pts * bytes_pp / ms / 1e6
    print(json.dumps(dict(config=config, case=name, ms=round(ms, 4), gpts_per_s=round(pts / ms / 1e6, 2), hbm_gbs=round(gbs, 1),
                          frac_of_measured_peak=round(gbs / PEAK, 3), frac_of_8TBs=round(gbs / 8000, 3), kernel=_cabi.last_kernel(), **extra)), flush=True)


if "1" in which:  # README divergence 100^3, accuracy 6 (L2-resident, launch-latency bound)
    F = torch.randn(3, 100, 100, 100, device=dev)
    h = 1.0 / 99
    ms = timeit(lambda: fdx.divergence(F, step_size=(h, h, h), keepdims=False, accuracy=6, method="central"), reps=200, warm=20)
    report(1, "divergence (3,100^3) fp32 a=6 (back-to-back calls, includes Python launch path)", 100**3, 16, ms, latency_us=round(ms * 1e3, 1))

if "2" in which:  # difference sweep on 8192^2
    x = torch.randn(8192, 8192, device=dev)
    res = []
    for axis in (0, 1):
        for method in ("central", "forward", "backward"):
            for d in (1, 2, 3, 4):
                for a in range(2, 9):
                    ms = timeit(lambda: fdx.difference(x, axis=axis, accuracy=a, derivative=d, method=method, step_size=0.5), reps=5, warm=1)
                    res.append((ms, axis, method, d, a, _cabi.last_kernel()))
    for axis in (0, 1):
        sub = sorted(r[0] for r in res if r[1] == axis)
        med = sub[len(sub) // 2]
        report(2, f"difference 8192^2 fp32 axis={axis}: median of {len(sub)} (d 1-4 x a 2-8 x 3 methods)", 8192**2, 8, med,
               best_ms=round(sub[0], 4), worst_ms=round(sub[-1], 4))
    worst = max(res)
    print(json.dumps(dict(config=2, slowest_case=dict(ms=round(worst[0], 4), axis=worst[1], method=worst[2], derivative=worst[3], accuracy=worst[4]))))

if "3" in which:  # laplacian / gradient 512^3 a=4 and their vjps
    n = 512
    h = 1.0 / (n - 1)
    xs = [torch.randn(n, n, n, device=dev) for _ in range(3)]
    it = iter(range(10**9))
    report(3, "laplacian 512^3 fp32 a=4", n**3, 8, timeit(lambda: fdx.laplacian(xs[next(it) % 3], accuracy=4, step_size=h)))
    report(3, "gradient 512^3 fp32 a=4", n**3, 16, timeit(lambda: fdx.gradient(xs[next(it) % 3], accuracy=4, step_size=h)))
    x = xs[0].requires_grad_(True)
    out = fdx.laplacian(x, accuracy=4, step_size=h)
    g = torch.randn_like(out)
    report(3, "vjp of laplacian 512^3 fp32 a=4 (transposed plans)", n**3, 8, timeit(lambda: torch.autograd.grad(out, x, g, retain_graph=True)))
    outg = fdx.gradient(x, accuracy=4, step_size=h)
    gg = torch.randn_like(outg)
    report(3, "vjp of gradient 512^3 fp32 a=4 (= fused divergence with transposed plans)", n**3, 16, timeit(lambda: torch.autograd.grad(outg, x, gg, retain_graph=True)))
    del xs, x, out, g, outg, gg
    torch.cuda.empty_cache()

if "4" in which:  # divergence / curl on (3,1024^3) fp32, accuracy 4 (assumed, not named in BASELINE.json)
    n = 1024
    h = 1.0 / (n - 1)
    F = torch.randn(3, n, n, n, device=dev)
    report(4, "divergence (3,1024^3) fp32 a=4", n**3, 16, timeit(lambda: fdx.divergence(F, accuracy=4, step_size=h, keepdims=False), reps=5, warm=2))
    report(4, "curl (3,1024^3) fp32 a=4", n**3, 24, timeit(lambda: fdx.curl(F, accuracy=4, step_size=h), reps=5, warm=2))
    del F
    torch.cuda.empty_cache()

if "5" in which:  # laplacian fp64 a=8: the slab one of 8 GPUs owns in the 2048^3 problem (256 planes + halos)
    n = 2048
    h = 1.0 / (n - 1)
    x = torch.randn(256 + 10, n, n, device=dev, dtype=torch.float64)
    report(5, "laplacian fp64 a=8 on a (266,2048,2048) slab = 1/8 of 2048^3 incl. halos", x.numel(), 16,
           timeit(lambda: fdx.laplacian(x, accuracy=8, step_size=h), reps=5, warm=2))
    x1 = torch.randn(512, 512, 512, device=dev, dtype=torch.float64)
    report(5, "laplacian 512^3 fp64 a=8", 512**3, 16, timeit(lambda: fdx.laplacian(x1, accuracy=8, step_size=h), reps=10, warm=2))
