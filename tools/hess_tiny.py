"""Tiny and awkward shapes through the fused hessians against the composition (prints mismatches)."""
import itertools
import os

import torch

import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi

bad = n = nf = 0
for dtype, v in ((torch.float32, 4), (torch.float64, 2)):
    for method, acc in (("central", 2), ("central", 3), ("central", 4), ("forward", 1), ("forward", 2), ("backward", 2), ("backward", 3)):
        dims0 = [4, 5, 6, 7, 8, 9, 10, 13, 17]
        dimsx = [2 * v, 3 * v, 4 * v, 5 * v, 16 * v, 17 * v, 32 * v + v, 33 * v]
        shapes = [(a, b, c) for a, b, c in itertools.product(dims0[::2], dims0[1::2], dimsx)] + [(a, c) for a, c in itertools.product(dims0, dimsx)]
        for shape in shapes:
            x = torch.randn(shape, device="cuda", dtype=dtype)
            kw = dict(accuracy=acc, step_size=0.37, method=method)
            try:
                h = fdx.hessian(x, **kw)
            except ValueError:
                continue
            k = _cabi.last_kernel()
            os.environ["FDX_HESS_DISABLE"] = "1"
            _cabi.reload_env()
            h0 = fdx.hessian(x, **kw)
            os.environ.pop("FDX_HESS_DISABLE")
            _cabi.reload_env()
            n += 1
            if not k.startswith("fused_hessian"):
                continue
            nf += 1
            sc = h0.abs().amax().clamp_min(1e-30)
            err = ((h - h0).abs().amax() / sc).item()
            if not err <= (5e-5 if dtype == torch.float32 else 1e-11):
                bad += 1
                print("MISMATCH", dtype, method, acc, shape, k, err, flush=True)
print(f"{n} cases, {nf} fused, {bad} mismatches")
