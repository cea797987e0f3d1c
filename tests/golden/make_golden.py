"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own source.

Usage (in the build container only; /root/reference does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (``/root/reference/finitediffx``) is imported unmodified.  Because ``jax`` is not
installed in this image, ``_jax_numpy_standin`` is registered as ``jax`` first (see that file
for what it does and does not emulate).  Outputs:

    tests/golden/golden_x64.npz    float64 cases (the mode all of the reference's tests use)
    tests/golden/golden_x32.npz    float32 cases, jax_enable_x64 off (int32/float32 coefficient solve)
    tests/golden/manifest.json     the call behind every array in the two files

Each case stores its input, its keyword arguments and the reference's output, so the tests need
neither the reference nor this script at run time.
"""

from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

import _jax_numpy_standin as standin  # noqa: E402

standin.install()
import finitediffx as ref  # noqa: E402  (the reference, unmodified)


def _cases():
    rng = np.random.default_rng(20261019)
    cases = []

    def add(op, x, **kw):
        cases.append((op, np.asarray(x), kw))

    # --- the reference's own golden inputs (tests/test_finite_diff.py:40-58, 76-108, 126-158)
    for method in ("central", "forward", "backward"):
        add("difference", [1.0, 4, 6, 3, 2], accuracy=1, axis=0, derivative=1, method=method)
        add("difference", [1.0, 4, 6, 3, 2, 4], accuracy=2, axis=0, derivative=1, method=method)
        add("difference", [1.2, 1.4], accuracy=1, axis=0, derivative=1, method=method)
    add("difference", [1.2, 1.3, 2.2, 3.0, 4.5, 5.5, 6.0, 7.0, 8.0, 20.0], accuracy=1)  # finite_diff.py:206-208

    # --- 1-D sweep: every (derivative, accuracy, method) of BASELINE config 2 and beyond
    x1 = rng.standard_normal(40)
    for d in (1, 2, 3, 4):
        for a in range(1, 9):
            for method in ("central", "forward", "backward"):
                add("difference", x1, accuracy=a, axis=0, derivative=d, method=method, step_size=0.37)

    # --- minimal legal sizes (n == p+L for central, n == 2L one-sided)
    for d, a in ((1, 1), (1, 2), (2, 2), (1, 4), (2, 4), (3, 3)):
        p, L = (d + a) // 2, d + a - 1
        add("difference", rng.standard_normal(max(p + L, 2 * p + 1)), accuracy=a, derivative=d, method="central")
        add("difference", rng.standard_normal(2 * L), accuracy=a, derivative=d, method="forward")
        add("difference", rng.standard_normal(2 * L), accuracy=a, derivative=d, method="backward")

    # --- N-D, every axis incl. negative
    x2 = rng.standard_normal((13, 17))
    x3 = rng.standard_normal((10, 11, 12))
    for axis in (0, 1, -1):
        for method in ("central", "forward", "backward"):
            add("difference", x2, accuracy=3, axis=axis, derivative=2, method=method, step_size=0.25)
    for axis in (0, 1, 2, -2):
        add("difference", x3, accuracy=4, axis=axis, derivative=1, method="central", step_size=1.0 / 7)
        add("difference", x3, accuracy=2, axis=axis, derivative=3, method="forward", step_size=2.0)

    # --- vector-calculus operators
    f3 = rng.standard_normal((3, 12, 13, 14))
    f2 = rng.standard_normal((2, 11, 13))
    f4 = rng.standard_normal((4, 9, 10, 11))
    for method in ("central", "forward", "backward"):
        add("gradient", x3, accuracy=4, step_size=(0.1, 0.2, 0.3), method=method)
        add("gradient", x2, accuracy=(2, 5), step_size=0.5, method=method)
        add("laplacian", x3, accuracy=4, step_size=1.0 / 11, method=method)
        add("laplacian", x2, accuracy=(3, 6), step_size=(0.5, 0.25), method=method)
        add("divergence", f3, accuracy=6, step_size=(1.0 / 11, 1.0 / 12, 1.0 / 13), keepdims=False, method=method)
        add("divergence", f2, accuracy=3, step_size=(0.1, 0.3), keepdims=True, method=method)
        add("curl", f3, accuracy=4, step_size=(0.1, 0.2, 0.3), method=method)
        add("curl", f2, accuracy=2, step_size=0.2, method=method, keepdims=False)
        add("curl", f2, accuracy=2, step_size=0.2, method=method, keepdims=True)
        add("jacobian", f2, accuracy=4, step_size=(0.1, 0.2), method=method)
        add("jacobian", f4, accuracy=3, step_size=(0.1, 0.2, 0.3), method=method)
        add("hessian", x2, accuracy=4, step_size=(0.1, 0.2), method=method)
        add("hessian", x3, accuracy=2, step_size=(0.1, 0.2, 0.3), method=method)  # F4: key collision, bug-compatible
    add("laplacian", x3, accuracy=8, step_size=0.01, method="central")  # axis too short: lax.slice error
    add("laplacian", rng.standard_normal((14, 15, 16)), accuracy=8, step_size=0.01, method="central")
    add("divergence", f4, accuracy=2, step_size=0.5, keepdims=True, method="central")  # 4th component ignored
    add("gradient", x1, accuracy=2)  # defaults: int step_size stays an int
    add("laplacian", x3)  # defaults (accuracy=1)
    add("hessian", x2)  # default accuracy=2

    # --- README example at reduced size (README.md:43-70): sin/cos field, accuracy=6
    n = 16
    g = np.linspace(0, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    F = np.stack([np.sin(X) * np.cos(Y), np.cos(Y) * Z**2, np.sin(Z) * X], axis=0)
    h = float(g[1] - g[0])
    add("divergence", F, accuracy=6, step_size=(h, h, h), keepdims=False, method="central")
    add("curl", F, accuracy=6, step_size=(h, h, h), method="central")
    add("laplacian", F[0], accuracy=6, step_size=(h, h, h), method="central")
    return cases


def _jsonable(kw):
    return {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}


def _run(x64: bool):
    standin.set_x64(x64)
    dt = np.float64 if x64 else np.float32
    arrays, manifest = {}, []
    for i, (op, x, kw) in enumerate(_cases()):
        xin = x.astype(dt)
        try:
            out = np.asarray(getattr(ref, op)(xin.view(standin.Array), **kw))
            err = None
        except Exception as e:  # the reference's own failure modes are part of its behaviour
            out, err = None, f"{type(e).__name__}"
        key = "in_" + hashlib.sha1(xin.tobytes() + str(xin.shape).encode()).hexdigest()[:12]
        arrays[key] = xin
        if out is not None:
            assert out.dtype == dt, (op, kw, out.dtype)
            arrays[f"out_{i}"] = out
        manifest.append({"id": i, "input": key, "op": op, "kwargs": _jsonable(kw), "error": err, "shape": list(x.shape)})
    return arrays, manifest


def _coeff_tables():
    """Coefficients the reference itself produces, both widths, for every stencil of config 2."""
    from finitediffx._src import utils as U

    rows = {}
    for x64 in (True, False):
        standin.set_x64(x64)
        for d in (1, 2, 3, 4):
            for a in range(1, 9):
                for kind, offs in (
                    ("central", U._generate_central_offsets(d, a + 1)),
                    ("forward", U._generate_forward_offsets(d, a)),
                    ("backward", U._generate_backward_offsets(d, a)),
                ):
                    c = np.asarray(ref.generate_finitediff_coeffs(offs, d))
                    rows[f"{'x64' if x64 else 'x32'}_{kind}_d{d}_a{a}"] = c
    return rows


def main():
    a64, m64 = _run(True)
    a32, m32 = _run(False)
    np.savez_compressed(os.path.join(HERE, "golden_x64.npz"), **a64)
    np.savez_compressed(os.path.join(HERE, "golden_x32.npz"), **a32)
    np.savez_compressed(os.path.join(HERE, "golden_coeffs.npz"), **_coeff_tables())
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(
            {
                "generator": "tests/golden/make_golden.py",
                "reference": "ASEM000/finitediffX v%s, source imported unmodified from /root/reference" % ref.__version__,
                "backend": "NumPy stand-in for jax (tests/golden/_jax_numpy_standin.py); XLA itself did not run",
                "x64": m64,
                "x32": m32,
            },
            f,
            indent=1,
        )
    nerr = sum(1 for m in m64 if m["error"])
    print(f"wrote {len(m64)} x64 cases ({nerr} reference errors), {len(m32)} x32 cases")
    for name in ("golden_x64.npz", "golden_x32.npz", "golden_coeffs.npz"):
        print(name, os.path.getsize(os.path.join(HERE, name)) // 1024, "KiB")


if __name__ == "__main__":
    main()
