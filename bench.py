#!/usr/bin/env python
"""bench.py -- finitediffx_b200 headline benchmark (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one synthetic grid.  Headline workload (N=1):
fdx.laplacian on a 512^3 fp32 grid, accuracy=4, central (BASELINE config 3, the configuration the
north-star target is quoted on).  With N>1 the global grid is (512*N, 512, 512), slab-decomposed
along axis 0 with NCCL halo exchange: per-GPU work is fixed ("weak" scaling).

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "laplacian_gpts_per_s"
UNIT = "Gpts/s"

WORKLOADS = {
    # name: (op, spatial shape, n_in fields, n_out fields, dtype, accuracy, bytes per point)
    "laplacian_512_f32_a4": dict(op="laplacian", n=512, dtype="f32", accuracy=4, fields_in=1, fields_out=1),
    "divergence_512_f32_a4": dict(op="divergence", n=512, dtype="f32", accuracy=4, fields_in=3, fields_out=1),
    "gradient_512_f32_a4": dict(op="gradient", n=512, dtype="f32", accuracy=4, fields_in=1, fields_out=3),
    "curl_512_f32_a4": dict(op="curl", n=512, dtype="f32", accuracy=4, fields_in=3, fields_out=3),
    "laplacian_512_f64_a8": dict(op="laplacian", n=512, dtype="f64", accuracy=8, fields_in=1, fields_out=1),
    "laplacian_1024_f32_a4": dict(op="laplacian", n=1024, dtype="f32", accuracy=4, fields_in=1, fields_out=1),
    # BASELINE config 5: a FIXED 2048^3 fp64 grid, accuracy 8, slab-decomposed over the N ranks (strong scaling;
    # 137 GB of HBM at N=1, 17 GB per rank at N=8)
    "laplacian_2048_f64_a8_strong": dict(op="laplacian", n=2048, dtype="f64", accuracy=8, fields_in=1, fields_out=1, strong=True),
    # BASELINE config 4: a FIXED (3, 1024^3) fp32 field, slab-decomposed over the N ranks (strong scaling)
    "divergence_1024_f32_a4_strong": dict(op="divergence", n=1024, dtype="f32", accuracy=4, fields_in=3, fields_out=1, strong=True),
    "curl_1024_f32_a4_strong": dict(op="curl", n=1024, dtype="f32", accuracy=4, fields_in=3, fields_out=3, strong=True),
}
HEADLINE = "laplacian_512_f32_a4"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Polls NVML for SM clock and throttle reasons DURING the timed region."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {
            "sm_mhz": float(np.median(self.samples)),
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def _config(name, world):
    """The `config` object of the JSON line: a function of (workload, GPU count) only, so that both arms print the same."""
    w = WORKLOADS[name]
    n = w["n"]
    strong = bool(w.get("strong"))
    return {
        "workload": name,
        "grid": [n, n, n] if strong else [n * world, n, n],
        "grid_per_gpu": [n // world, n, n] if strong else [n, n, n],
        "accuracy": w["accuracy"],
        "method": "central",
        "step_size": f"1/{n - 1}",
        "decomposition": "single GPU" if world == 1 else f"slab along axis 0, {world} ranks, halo exchange over NVLink (see halo_transport)",
        "l2": "one input slab per rank, far larger than L2 (126 MB)" if strong else "3 rotating inputs, each larger than L2 (126 MB); outputs freshly written",
    }


def _bytes_per_point(w):
    s = 4 if w["dtype"] == "f32" else 8
    return s * (w["fields_in"] + w["fields_out"])


def _make_input(w, device, seed, n0=None):
    import torch

    n = w["n"]
    n0 = n if n0 is None else n0
    dt = torch.float32 if w["dtype"] == "f32" else torch.float64
    g = torch.Generator(device=device).manual_seed(seed)
    shape = (n0, n, n) if w["fields_in"] == 1 else (w["fields_in"], n0, n, n)
    return torch.randn(shape, generator=g, device=device, dtype=dt)


def _call(fdx, w, x):
    h = 1.0 / (w["n"] - 1)
    if w["op"] == "divergence":
        return fdx.divergence(x, accuracy=w["accuracy"], step_size=h, keepdims=False)
    return getattr(fdx, w["op"])(x, accuracy=w["accuracy"], step_size=h)


def _time_device(fn, inputs, steps, warmup, sampler=None, barrier=None):
    """W warm-up + K timed steps, CUDA events on the launching (current) stream."""
    import torch

    for i in range(warmup):
        fn(inputs[i % len(inputs)])
    torch.cuda.synchronize()
    if barrier:
        barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx = sampler if sampler is not None else _Null()
    with ctx:
        e0.record()
        for i in range(steps):
            fn(inputs[i % len(inputs)])
        e1.record()
        torch.cuda.synchronize()
    if barrier:
        barrier()
    return e0.elapsed_time(e1) / 1e3  # seconds


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ------------------------------------------------------------------------------ CPU arms
def _cpu_threads():
    import torch

    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return n


def _cpu_step(w, n):
    """One CPU step of workload `w` on an n^3 sample (restated reference graph on torch-CPU)."""
    import torch

    from oracle import fdx_oracle_torch as OT

    dt = torch.float32 if w["dtype"] == "f32" else torch.float64
    g = torch.Generator().manual_seed(0)
    shape = (n, n, n) if w["fields_in"] == 1 else (w["fields_in"], n, n, n)
    x = torch.randn(shape, generator=g, dtype=dt)
    h = 1.0 / (w["n"] - 1)

    def step():
        if w["op"] == "divergence":
            return OT.divergence(x, accuracy=w["accuracy"], step_size=h, keepdims=False)
        return getattr(OT, w["op"])(x, accuracy=w["accuracy"], step_size=h)

    return step


def _cpu_baseline(w, budget_s=12.0):
    cores = _cpu_threads()
    n = 256 if w["n"] >= 256 else w["n"]
    step = _cpu_step(w, n)
    step()  # warm-up
    t0, reps = time.perf_counter(), 0
    while True:
        step()
        reps += 1
        el = time.perf_counter() - t0
        if el >= budget_s or reps >= 50:
            break
    gpts = n**3 * reps / el / 1e9
    return {
        "value": gpts,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{w['op']} on a {n}^3 {w['dtype']} sub-grid of the workload, {reps} reps in {el:.1f}s; restated reference algorithm on torch-CPU, not jax[cpu] (jax is not installed)",
    }


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (torch-CPU restatement)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = WORKLOADS[args.workload]
    cores = _cpu_threads()
    # size the per-step sample so that (K+W) steps end within ~2 minutes
    n = 128
    step = _cpu_step(w, n)
    step()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter() - t0
    best = n
    for cand in (256, 512):
        if cand <= w["n"] and t1 * (cand / n) ** 3 * (args.steps + args.warmup) <= 120.0:
            best = cand
        else:
            break
    if best != n:
        n = best
        step = _cpu_step(w, n)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    value = n**3 * args.steps / el / 1e9
    sample = f"{w['op']} on a {n}^3 {w['dtype']} sub-grid per step; restated reference algorithm on torch-CPU, {cores} threads, not jax[cpu]"
    line = {
        "impl": "reference",
        "metric": METRIC if w["op"] == "laplacian" else f"{w['op']}_gpts_per_s",
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": w["dtype"],
        "data": "synthetic",
        "config": _config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch

    import finitediffx_b200 as fdx
    from finitediffx_b200 import _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    n = w["n"]
    pts_per_gpu = n**3
    _cabi.load()

    barrier = (lambda: dist.barrier()) if dist else None
    strong = bool(w.get("strong"))
    if strong:
        from finitediffx_b200 import distributed as fd

        if n % world:
            raise SystemExit(f"{args.workload}: {n} planes do not split over {world} ranks")
        pts_per_gpu = n**3 // world
        slab = fd.Slab(global_n0=n, rank=rank, world=world, group=dist.group.WORLD if dist else None)
        op = fd.SlabOperator(slab, w["op"], spatial=(n, n, n), dtype=torch.float32 if w["dtype"] == "f32" else torch.float64, accuracy=w["accuracy"], step_size=1.0 / (n - 1), device=dev, n_buffers=1)
        op.fill_random(0, seed=100 * rank)
        inputs = [0]
        fn = lambda b: op.step(b)
    elif world == 1:
        # rotate over 3 input buffers so nothing stays L2-resident between steps (inputs >> L2)
        inputs = [_make_input(w, dev, seed=s) for s in range(3)]
        fn = lambda x: _call(fdx, w, x)
    else:
        from finitediffx_b200 import distributed as fd

        slab = fd.Slab(global_n0=n * world, rank=rank, world=world, group=dist.group.WORLD)
        h = 1.0 / (n - 1)
        op = fd.SlabOperator(slab, w["op"], spatial=(n, n, n), dtype=torch.float32 if w["dtype"] == "f32" else torch.float64, accuracy=w["accuracy"], step_size=h, device=dev, n_buffers=3)
        inputs = list(range(3))
        for b in inputs:
            op.fill_random(b, seed=100 * rank + b)
        fn = lambda b: op.step(b)

    launches0 = _cabi.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # count launches over the timed region only
    for i in range(args.warmup):
        fn(inputs[i % len(inputs)])
    torch.cuda.synchronize()
    graphed = False
    halo_mode = op.mode if world > 1 or strong else None
    want_graph = os.environ.get("FDX_BENCH_GRAPH", "1" if halo_mode == "peer" else "0") == "1"
    if world > 1 and want_graph:
        # a multi-GPU step is a dozen enqueue calls (barriers, peer copies, events, two launches) for 0.2 ms of GPU work:
        # one step per rotating buffer is captured in a CUDA graph and replayed (same kernels, exchange and buffers).
        # Default with peer-memory halos (nothing but memcpys and kernels in the graph); opt-in with NCCL halos
        # (FDX_BENCH_GRAPH=1: tearing the process group down after an NCCL capture hung on this stack).
        for b in inputs:
            op.capture(b)
        fn = lambda b: op.replay(b)
        for i in range(3):
            fn(inputs[i % len(inputs)])
        torch.cuda.synchronize()
        graphed = True
    launches0 = _cabi.launch_count()
    secs = _time_device(fn, inputs, args.steps, 0, sampler=sampler, barrier=barrier)
    launches = _cabi.launch_count() - launches0
    if graphed:
        launches = op.launches_per_step * args.steps  # replayed launches are not seen by the library's counter
    if dist:
        t = torch.tensor([secs], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    value = pts_per_gpu * world * args.steps / secs / 1e9
    ms = secs / args.steps * 1e3

    # ---- end-to-end through the public API with HOST buffers (H2D + kernel + D2H every step)
    e2e = None
    if world == 1 and not strong:
        hx = [torch.empty(inputs[0].shape, dtype=inputs[0].dtype, pin_memory=True).copy_(inputs[i]) for i in range(2)]
        e_steps = max(3, min(args.steps, 10))
        r = None
        for i in range(3):  # warm-up as the timed loop runs: the previous result stays alive while the next
            r = _call(fdx, w, hx[i % 2])  # one is produced, so both pinned result blocks exist before timing
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e_steps):
            r = _call(fdx, w, hx[i % 2])
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        assert not r.is_cuda
        e2e = {
            "value": pts_per_gpu * e_steps / el / 1e9,
            "unit": UNIT,
            "h2d_bytes_per_step": int(hx[0].numel() * hx[0].element_size()),
            "d2h_bytes_per_step": int(r.numel() * r.element_size()),
            "steps": e_steps,
            "api": f"finitediffx_b200.{w['op']}(pinned host tensor) -> host tensor",
        }
    elif world > 1 and not strong:
        # every rank: its slab from pinned HOST memory to the device (H2D), one distributed step (halo exchange
        # included), the result back to pinned host memory (D2H); all ranks at once, each over its own PCIe link
        b0 = inputs[0]
        hin = [torch.empty(op.owned_view(b0, f).shape, dtype=op.dtype, pin_memory=True).copy_(op.owned_view(b0, f)) for f in range(len(op.inputs[b0]))]
        hout = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in op.outputs[b0]]

        def e2e_step():
            for f, h in enumerate(hin):
                op.owned_view(b0, f).copy_(h, non_blocking=True)
            outs_ = fn(b0)
            for h, t in zip(hout, outs_):
                h.copy_(t, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            e2e_step()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        el = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e = {
            "value": pts_per_gpu * world * e_steps / float(el.item()) / 1e9,
            "unit": UNIT,
            "h2d_bytes_per_step": int(sum(h.numel() * h.element_size() for h in hin)) * world,
            "d2h_bytes_per_step": int(sum(h.numel() * h.element_size() for h in hout)) * world,
            "steps": e_steps,
            "api": "SlabOperator: pinned host slab -> H2D -> step (halo exchange + kernels) -> D2H, every rank over its own PCIe link; bytes are totals over ranks",
        }
    else:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "the strong-scaling workloads keep their slabs device-resident; e2e is measured on the headline workload"}

    # ---- the other BASELINE configs, so that the driver's run holds their numbers too (all ranks take part at N>1)
    peak, peak_src = _peaks()
    extras = None
    if not args.no_extras and args.workload == HEADLINE:
        del inputs
        if world > 1:
            op.release_graphs()
            del op, fn
        torch.cuda.empty_cache()
        try:
            extras = _extras_single(fdx, _cabi, dev, peak) if world == 1 else _extras_multi(world, rank, dev, dist, peak)
        except Exception as e:  # the headline line must not be lost to an extra (e.g. a smaller-memory GPU)
            extras = {"error": repr(e)[:300]}

    if rank != 0:
        if dist:
            _teardown(dist, graphed and halo_mode != "peer")
        return 0

    alg_bytes = _bytes_per_point(w) * pts_per_gpu
    achieved = alg_bytes / (secs / args.steps) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and os.path.exists(tpath):  # per-launch DRAM bytes of the single-GPU kernel from its ncu --set full capture
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get(args.workload), tj.get("_source")
        except Exception:
            traffic = None
    config = _config(args.workload, world)
    line = {
        "metric": METRIC if w["op"] == "laplacian" else f"{w['op']}_gpts_per_s",
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None,
        "dtype": w["dtype"],
        "data": "synthetic",
        "config": config,
        "graph_replay": graphed,
        "halo_transport": halo_mode,
        "clocks": sampler.summary() if sampler else None,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "kernel": "fused_" + w["op"] + "3d",
        "roofline": {
            "bound": "hbm",
            "achieved": achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved / peak,
            "frac_of_8TBs_nominal": achieved / 8000.0,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes,
            "traffic": traffic,
            "traffic_source": traffic_src,
        },
    }
    if world == 1 and not args.no_cpu and not strong:
        line["cpu_baseline"] = _cpu_baseline(w)
    if extras is not None:
        line["extras"] = extras
    print(json.dumps(line), flush=True)
    if dist:
        _teardown(dist, graphed and halo_mode != "peer")
    return 0


def _teardown(dist, nccl_captured):
    import torch

    torch.cuda.synchronize()
    if nccl_captured:
        # (opt-in mode) tearing the process group down with captured NCCL work hung on this stack
        # (torch 2.11 / NCCL 2.28.9): leave without the teardown
        sys.stdout.flush()
        os._exit(0)
    dist.destroy_process_group()


# ------------------------------------------------------------------------------ extras: every BASELINE config
def _entry(pts, bytes_per_pt, secs_per_step, peak, kernel=None, **more):
    gbs = pts * bytes_per_pt / secs_per_step / 1e9
    e = {"ms": secs_per_step * 1e3, "gpts_per_s": pts / secs_per_step / 1e9, "hbm_gbs": gbs, "frac": gbs / peak, "frac_of_8TBs_nominal": gbs / 8000.0}
    if kernel:
        e["kernel"] = kernel
    e.update(more)
    return e


def _extras_single(fdx, _cabi, dev, peak):
    """Configs 1-5 of BASELINE.json on one GPU: device-resident, CUDA events, 5 warm-up + 20 timed steps each (inputs
    far larger than L2, or three rotating inputs).  `frac` = algorithmic bytes / time / measured HBM peak."""
    import torch

    ex = {"note": "device-resident, CUDA events on the launching stream, 5 warm-up + 20 timed calls per case; frac = algorithmic bytes / time / measured HBM copy peak"}
    K, W = 20, 5

    def timed(fn, inputs, k=K, wu=W):
        return _time_device(fn, inputs, k, wu) / k

    # ---- config 3 (and the divergence half of the north-star target): 512^3 fp32, accuracy 4, forward and vjp
    for name in ("divergence_512_f32_a4", "gradient_512_f32_a4", "curl_512_f32_a4", "laplacian_512_f64_a8"):
        we = WORKLOADS[name]
        inputs = [_make_input(we, dev, seed=s) for s in range(3)]
        t = timed(lambda x: _call(fdx, we, x), inputs)
        ex[name] = _entry(we["n"] ** 3, _bytes_per_point(we), t, peak, _cabi.last_kernel())
        del inputs
    for name in ("laplacian_512_f32_a4", "gradient_512_f32_a4", "divergence_512_f32_a4", "curl_512_f32_a4"):
        we = WORKLOADS[name]
        x = _make_input(we, dev, seed=0).requires_grad_(True)
        out = _call(fdx, we, x)
        gs = [torch.randn_like(out) for _ in range(3)]
        t = timed(lambda g: torch.autograd.grad(out, x, g, retain_graph=True), gs)
        ex[name + "_vjp"] = _entry(we["n"] ** 3, _bytes_per_point(we), t, peak, _cabi.last_kernel())
        if name == HEADLINE:
            def fwd_bwd(g):
                o = _call(fdx, we, x)
                return torch.autograd.grad(o, x, g)

            t = timed(fwd_bwd, gs)
            ex[name + "_fwd_plus_vjp"] = _entry(we["n"] ** 3, 2 * _bytes_per_point(we), t, peak)
        del x, out, gs
    # ---- the headline again: 200 steps back to back, and ~2 s of sustained load (SM clocks settle under power)
    we = WORKLOADS[HEADLINE]
    inputs = [_make_input(we, dev, seed=s) for s in range(3)]
    t = timed(lambda x: _call(fdx, we, x), inputs, k=200, wu=20)
    ex["laplacian_512_f32_a4_200_steps"] = _entry(we["n"] ** 3, _bytes_per_point(we), t, peak)
    samp = ClockSampler(dev.index or 0)
    t_all = _time_device(lambda x: _call(fdx, we, x), inputs, 8000, 2000, sampler=samp)
    ex["laplacian_512_f32_a4_sustained"] = _entry(we["n"] ** 3, _bytes_per_point(we), t_all / 8000, peak, steps=8000, warmup=2000, clocks=samp.summary())
    del inputs
    torch.cuda.empty_cache()
    # ---- config 1: README divergence (3, 100^3) fp32 accuracy 6: L2-resident, latency per call (host launch path included)
    g = torch.linspace(0, 1, 100, device=dev)
    X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
    F = torch.stack([torch.sin(X) * torch.cos(Y), torch.cos(Y) * Z**2, torch.sin(Z) * X])
    h = 1.0 / 99
    t = timed(lambda f: fdx.divergence(f, accuracy=6, step_size=(h, h, h), keepdims=False), [F], k=200, wu=20)
    ex["config1_divergence_100_f32_a6"] = {"us_per_call": t * 1e6, "gpts_per_s": 1e6 / t / 1e9, "kernel": _cabi.last_kernel(),
                                           "note": "16 MB working set: L2-resident and launch-latency bound; back-to-back calls through the public API"}
    del F, X, Y, Z
    # ---- config 2: difference on (8192, 8192) fp32, all 168 cases (derivative 1-4, accuracy 2-8, three methods, both axes)
    x = torch.randn((8192, 8192), device=dev)
    per_axis = {0: [], 1: []}
    for method in ("central", "forward", "backward"):
        for d in range(1, 5):
            for a in range(2, 9):
                for axis in (0, 1):
                    # (best of two runs of 10 calls: a host hiccup inside a 1 ms timed region would otherwise pass for a slow case)
                    call = lambda v: fdx.difference(v, axis=axis, accuracy=a, derivative=d, method=method, step_size=0.5)  # noqa: E731
                    t = min(timed(call, [x], k=10, wu=3), timed(call, [x], k=10, wu=1))
                    per_axis[axis].append((t, method, d, a))
    pts = 8192 * 8192
    for axis, rows in per_axis.items():
        rows.sort()
        med, worst, best = rows[len(rows) // 2], rows[-1], rows[0]
        ex[f"config2_difference_8192sq_f32_axis{axis}"] = {
            "cases": len(rows),
            "median": _entry(pts, 8, med[0], peak),
            "best": _entry(pts, 8, best[0], peak, case=list(best[1:])),
            "worst": _entry(pts, 8, worst[0], peak, case=list(worst[1:])),
            "note": "per call through the public API (host launch path included); per case the better of two runs of 10 calls",
        }
    # ---- beyond the BASELINE configs: the 2-D operators on the same 8192^2 grid (most of the reference's own tests are 2-D)
    f2 = torch.randn((2, 8192, 8192), device=dev)
    for name, call, arr, fields in (("laplacian2d_8192sq_f32_a4", lambda v: fdx.laplacian(v, accuracy=4, step_size=0.1), x, 2),
                                    ("gradient2d_8192sq_f32_a4", lambda v: fdx.gradient(v, accuracy=4, step_size=0.1), x, 3),
                                    ("divergence2d_8192sq_f32_a4", lambda v: fdx.divergence(v, accuracy=4, step_size=0.1), f2, 3),
                                    ("curl2d_8192sq_f32_a4", lambda v: fdx.curl(v, accuracy=4, step_size=0.1), f2, 3)):
        t = timed(call, [arr], k=10, wu=3)
        ex[name] = _entry(pts, 4 * fields, t, peak, _cabi.last_kernel())
    t = timed(lambda v: fdx.hessian(v, accuracy=4, step_size=0.1), [x], k=10, wu=3)
    ex["hessian2d_8192sq_f32_a4"] = _entry(pts, 4 * (1 + 4), t, peak, _cabi.last_kernel(), note="ONE fused launch (fdx_hessian2d); one read + four writes")
    del x, f2
    torch.cuda.empty_cache()
    # ---- jacobian and hessian on 3-D grids (algorithmic bytes: one read per input field, one write per output field)
    f3 = torch.randn((3, 384, 384, 384), device=dev)
    t = timed(lambda v: fdx.jacobian(v, accuracy=4, step_size=0.1), [f3], k=10, wu=3)
    ex["jacobian_3x384_f32_a4"] = _entry(384**3, 4 * (3 + 9), t, peak, _cabi.last_kernel(), note="one fused gradient per component")
    t = timed(lambda v: fdx.hessian(v, accuracy=4, step_size=0.1), [f3[0]], k=10, wu=3)
    ex["hessian_384_f32_a4"] = _entry(384**3, 4 * (1 + 9), t, peak, _cabi.last_kernel(),
                                      note="ONE fused launch (fdx_hessian3d: the reference's table entries, bug-compatible, stored in all nine slots); one read + nine writes")
    x64 = f3[0].double()
    del f3
    t = timed(lambda v: fdx.hessian(v, accuracy=4, step_size=0.1), [x64], k=10, wu=3)
    ex["hessian_384_f64_a4"] = _entry(384**3, 8 * (1 + 9), t, peak, _cabi.last_kernel())
    del x64
    torch.cuda.empty_cache()
    # ---- config 4: divergence / curl on a (3, 1024^3) fp32 field, accuracy 4 (one input: 12.9 GB, far larger than L2)
    for name in ("divergence_1024_f32_a4_strong", "curl_1024_f32_a4_strong"):
        we = WORKLOADS[name]
        x = torch.empty((3, 1024, 1024, 1024), device=dev)
        for c in range(3):
            x[c].normal_()
        t = timed(lambda v: _call(fdx, we, v), [x], k=10, wu=3)
        ex["config4_" + name.replace("_strong", "")] = _entry(1024**3, _bytes_per_point(we), t, peak, _cabi.last_kernel())
        del x
        torch.cuda.empty_cache()
    # ---- config 5: laplacian fp64 accuracy 8; one rank's slab of the 2048^3 grid at 8 GPUs (256 planes + 2 x 5 halo planes)
    we = WORKLOADS["laplacian_2048_f64_a8_strong"]
    x = torch.empty((266, 2048, 2048), device=dev, dtype=torch.float64)
    x.normal_()
    t = timed(lambda v: fdx.laplacian(v, accuracy=8, step_size=1.0 / 2047), [x], k=10, wu=3)
    ex["config5_laplacian_f64_a8_slab_266x2048x2048"] = _entry(266 * 2048 * 2048, 16, t, peak, _cabi.last_kernel())
    del x
    torch.cuda.empty_cache()
    return ex


def _extras_multi(world, rank, dev, dist, peak):
    """Strong scaling of BASELINE configs 4 and 5 over the N ranks of this run (device time, max over ranks), and -- the
    honest overlap figure -- the same slab computed by rank 0 alone with no exchange (`slab_alone_ms`)."""
    import torch

    from finitediffx_b200 import distributed as fd

    ex = {"note": "fixed global grids, slab-decomposed over the ranks of this run; ms = max over ranks, CUDA events, 3 warm-up + 10 timed steps; "
                  "slab_alone_ms = this rank's planes (with halos in place) through the same kernel without any exchange"}
    free_gb = torch.cuda.mem_get_info(dev)[0] / 1e9
    for name in ("divergence_1024_f32_a4_strong", "curl_1024_f32_a4_strong", "laplacian_2048_f64_a8_strong"):
        w = WORKLOADS[name]
        n = w["n"]
        esz = 4 if w["dtype"] == "f32" else 8
        need_gb = (w["fields_in"] + w["fields_out"]) * esz * n**3 / world / 1e9 * 1.1
        if n % world or need_gb > free_gb:
            ex[name] = {"skipped": f"needs {need_gb:.0f} GB per rank"}
            continue
        dt = torch.float32 if w["dtype"] == "f32" else torch.float64
        slab = fd.Slab(global_n0=n, rank=rank, world=world, group=dist.group.WORLD)
        op = fd.SlabOperator(slab, w["op"], spatial=(n, n, n), dtype=dt, accuracy=w["accuracy"], step_size=1.0 / (n - 1), device=dev, n_buffers=1)
        op.fill_random(0, seed=100 * rank)
        step = lambda b: op.step(b)
        if op.mode == "peer" and os.environ.get("FDX_BENCH_GRAPH", "1") == "1":
            op.step(0)
            op.capture(0)
            step = lambda b: op.replay(b)
        secs = _time_device(step, [0], 10, 3, barrier=lambda: dist.barrier())
        t = torch.tensor([secs], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_s = float(t.item()) / 10
        alone = _time_device(lambda b: op.step_alone(b), [0], 10, 3) / 10
        dist.barrier()
        e = _entry(n**3, _bytes_per_point(w), step_s, peak * world, n_gpus=world, slab_alone_ms=alone * 1e3, overlap_efficiency=alone / step_s,
                   halo_bytes_per_step_per_rank=int(op.halo_bytes_per_step), halo_transport=op.mode)
        ex[name] = e
        op.release_graphs()
        del op, step
        torch.cuda.empty_cache()
    return ex


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip the other BASELINE configs (extras)")
    ap.add_argument("--extras", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
