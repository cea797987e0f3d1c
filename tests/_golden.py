"""Loader for the committed golden vectors (tests/golden/, produced by make_golden.py)."""
import json
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(mode):
    """mode: 'x64' or 'x32' -> list of dict(id, op, kwargs, x, out|None, error)."""
    with open(os.path.join(HERE, "manifest.json")) as f:
        manifest = json.load(f)[mode]
    data = np.load(os.path.join(HERE, f"golden_{mode}.npz"))
    cases = []
    for m in manifest:
        kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in m["kwargs"].items()}
        out = data[f"out_{m['id']}"] if m["error"] is None else None
        cases.append(dict(id=m["id"], op=m["op"], kwargs=kw, x=data[m["input"]], out=out, error=m["error"]))
    return cases


def coeffs():
    return np.load(os.path.join(HERE, "golden_coeffs.npz"))
