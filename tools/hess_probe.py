"""Time the one-pass 3-D hessian (fdx_hessian3d) against the 7-pass composition and sweep its chunking switches.
usage: python tools/hess_probe.py [n] [accuracy]"""
import os
import sys

import torch

import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi


def timed(fn, k=10, wu=3):
    for _ in range(wu):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
    acc = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    for dt in (torch.float32, torch.float64):
        x = torch.randn((n, n, n), device="cuda", dtype=dt)
        s = x.element_size()
        gb = 10 * s * n**3 / 1e9
        for env in ({"FDX_HESS_DISABLE": "1"}, {}, {"FDX_HESS_ZC": "24"}, {"FDX_HESS_ZC": "48"}, {"FDX_HESS_ZC": "64"}):
            for k in ("FDX_HESS_DISABLE", "FDX_HESS_CPS", "FDX_HESS_ZC", "FDX_HESS_PART", "FDX_HESS_MINB", "FDX_HESS_PLAIN_ST", "FDX_HESS_XALIGN_B"):
                os.environ.pop(k, None)
            os.environ.update(env)
            _cabi.reload_env()
            med, best = timed(lambda: fdx.hessian(x, accuracy=acc, step_size=0.1))
            print(f"{dt} n={n} a={acc} {env or 'default'}: {_cabi.last_kernel()} median {med:.4f} ms  best {best:.4f} ms  {gb / med * 1e3:.0f} GB/s algorithmic", flush=True)
        del x


if __name__ == "__main__":
    main()
