"""Round-robin timing of a fused operator under experiment switches (env), with SM clocks.

    python tools/time_variants.py [op] [acc] [f32|f64] [n] -- "name:K=V,K=V" ...
Variants run interleaved for several rounds so that clock drift hits all of them alike; min and median are reported.
"""
import os, sys, json, statistics
os.environ["FDX_FUSED_EXPERIMENTS"] = "1"  # lets FDX_FUSED_DBG through (timing experiments: wrong results at the rim)
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import _cabi
try:
    import pynvml
    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(0)
    def clock():
        return pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
except Exception:
    def clock():
        return -1
args = sys.argv[1:]
specs = []
if "--" in args:
    i = args.index("--")
    specs = args[i + 1:]
    args = args[:i]
op = args[0] if len(args) > 0 else "laplacian"
acc = int(args[1]) if len(args) > 1 else 4
dt = torch.float64 if (len(args) > 2 and args[2] == "f64") else torch.float32
n = int(args[3]) if len(args) > 3 else 512
vjp = op.endswith("_vjp")
if vjp:
    op = op[:-4]
fin = 3 if op in ("divergence", "curl") else 1
shape = (n, n, n) if fin == 1 else (3, n, n, n)
xs = [torch.randn(shape, device="cuda", dtype=dt) for _ in range(3)]
h = 1.0 / (n - 1)
def fwd(x):
    if op == "divergence":
        return fdx.divergence(x, accuracy=acc, step_size=h, keepdims=False)
    return getattr(fdx, op)(x, accuracy=acc, step_size=h)
if vjp:
    xr = xs[0].requires_grad_(True)
    out0 = fwd(xr)
    gs = [torch.randn_like(out0) for _ in range(3)]
    it = iter(range(10**9))
    def call(x):
        return torch.autograd.grad(out0, xr, gs[next(it) % 3], retain_graph=True)
else:
    call = fwd
variants = [("default", {})]
for sp in specs:
    name, _, kv = sp.partition(":")
    env = dict(item.split("=") for item in kv.split(",") if item)
    variants.append((name, env))
keys = set(k for _, e in variants for k in e)
times = {name: [] for name, _ in variants}
clocks = {name: [] for name, _ in variants}
rounds = int(os.environ.get("ROUNDS", "7"))
for rnd in range(rounds):
    for name, env in variants:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        _cabi.reload_env()  # the library snapshots the FDX_* switches
        for i in range(3):
            call(xs[i % 3])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            call(xs[i % 3])
        e1.record()
        c = clock()
        torch.cuda.synchronize()
        times[name].append(e0.elapsed_time(e1) / 20)
        clocks[name].append(c)
for name, _ in variants:
    t = times[name]
    print(json.dumps(dict(variant=name, ms_min=round(min(t), 4), ms_med=round(statistics.median(t), 4), gpts_med=round(n**3 / statistics.median(t) / 1e6, 1),
                          sm_mhz=clocks[name])), flush=True)
