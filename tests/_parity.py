"""Parity metric of BASELINE.md section 5:  max_i |out_i - ref_i| / S_i  with
S_i = sum over the operator's terms of sum_k |c_k x_{i+k}| / h^d  (the conditioning scale)."""
import numpy as np

from oracle import fdx_oracle as O

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def _per_axis(v, nd):
    if isinstance(v, (tuple, list)):
        return tuple(v)
    return (v,) * nd


def scale_for(op, x, kwargs):
    """Conditioning scale of `op(x, **kwargs)`, broadcastable against the op's output."""
    x = np.asarray(x, dtype=np.float64)
    method = kwargs.get("method", "central")
    cs = lambda arr, ax, a, d, h: O.conditioning_scale(arr, axis=ax, accuracy=a, derivative=d, method=method, step_size=h)
    if op == "difference":
        return cs(x, kwargs.get("axis", 0), kwargs.get("accuracy", 1), kwargs.get("derivative", 1), kwargs.get("step_size", 1.0))
    if op == "laplacian":
        nd = x.ndim
        acc, hs = _per_axis(kwargs.get("accuracy", 1), nd), _per_axis(kwargs.get("step_size", 1), nd)
        return sum(cs(x, k, acc[k], 2, hs[k]) for k in range(nd))
    if op == "gradient":
        nd = x.ndim
        acc, hs = _per_axis(kwargs.get("accuracy", 1), nd), _per_axis(kwargs.get("step_size", 1), nd)
        return np.stack([cs(x, k, acc[k], 1, hs[k]) for k in range(nd)])
    if op == "jacobian":
        nd = x.ndim - 1
        acc, hs = _per_axis(kwargs.get("accuracy", 1), nd), _per_axis(kwargs.get("step_size", 1), nd)
        return np.stack([np.stack([cs(xi, k, acc[k], 1, hs[k]) for k in range(nd)]) for xi in x])
    if op == "divergence":
        nd = x.ndim - 1
        acc, hs = _per_axis(kwargs.get("accuracy", 1), nd), _per_axis(kwargs.get("step_size", 1), nd)
        s = sum(cs(x[k], k, acc[k], 1, hs[k]) for k in range(nd))
        return s[None] if kwargs.get("keepdims", True) else s
    if op == "curl":
        nd = x.ndim - 1
        acc, hs = _per_axis(kwargs.get("accuracy", 1), nd), _per_axis(kwargs.get("step_size", 1), nd)
        d = lambda c, ax: cs(x[c], ax, acc[ax], 1, hs[ax])
        if nd == 3:
            return np.stack([d(2, 1) + d(1, 2), d(0, 2) + d(2, 0), d(1, 0) + d(0, 1)])
        s = d(1, 0) + d(0, 1)
        return s[None] if kwargs.get("keepdims", True) else s
    if op == "hessian":
        nd = x.ndim
        acc, hs = _per_axis(kwargs.get("accuracy", 2), nd), _per_axis(kwargs.get("step_size", 1), nd)
        table = {}
        for k in range(nd):
            table[2 * k] = cs(x, k, acc[k], 2, hs[k])
        for a1 in range(nd):
            for a2 in range(a1 + 1, nd):
                inner = O.difference(x, axis=a1, accuracy=acc[a1], step_size=hs[a1], method=method)
                # |D2| |D1 x| plus the amplification of D1's own rounding by D2
                table[a1 + a2] = cs(np.abs(inner), a2, acc[a2], 1, hs[a2]) + cs(
                    cs(x, a1, acc[a1], 1, hs[a1]), a2, acc[a2], 1, hs[a2]
                )
        return np.stack([np.stack([table[i + j] for j in range(nd)]) for i in range(nd)])
    raise ValueError(op)


def parity_error(op, x, kwargs, out, ref):
    s = scale_for(op, x, kwargs)
    s = np.where(s == 0, 1.0, s)
    return float(np.max(np.abs(np.asarray(out, dtype=np.float64) - np.asarray(ref, dtype=np.float64)) / s))
