"""Time the one-pass 2-D hessian (fdx_hessian2d) against the composition: python tools/hess2d_probe.py [n] [accuracy]"""
import os
import sys

import torch

import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi
from hess_probe import timed


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    acc = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    for dt in (torch.float32, torch.float64):
        x = torch.randn((n, n), device="cuda", dtype=dt)
        gb = 5 * x.element_size() * n * n / 1e9
        for env in ({"FDX_HESS_DISABLE": "1"}, {"FDX_HESS2D_PF": "0"}, {"FDX_HESS2D_PF": "2"}, {"FDX_HESS2D_PF": "4"}, {"FDX_HESS2D_PF": "8"}, {"FDX_HESS2D_PF": "16"}):
            for k in ("FDX_HESS_DISABLE", "FDX_HESS2D_WPS", "FDX_HESS2D_MINB", "FDX_HESS2D_PF"):
                os.environ.pop(k, None)
            os.environ.update(env)
            _cabi.reload_env()
            med, best = timed(lambda: fdx.hessian(x, accuracy=acc, step_size=0.1))
            print(f"{dt} n={n} a={acc} {env or 'default'}: {_cabi.last_kernel()} median {med:.4f} ms  best {best:.4f} ms  {gb / med * 1e3:.0f} GB/s algorithmic", flush=True)
        del x


if __name__ == "__main__":
    main()
