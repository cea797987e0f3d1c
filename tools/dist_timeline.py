"""Timeline of one multi-GPU step from CUDA events (run under torchrun; ncu / nsys cannot wrap a multi-rank command).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_timeline.py [workload]

For both halo transports (peer-memory copies / NCCL send-recv) it prints, per rank, when the halo landed, when the shell
launch and the interior launch finished and when the step was done, all relative to the start of the step (medians over
20 eager steps), the eager and the graph-replayed step time (max over ranks), and the time of the same planes computed
alone (no exchange).  The summary kept under profiles/ is this output.
"""
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finitediffx_b200 import distributed as fd  # noqa: E402

CASES = {
    "laplacian_512_f32_a4_weak": dict(kind="laplacian", n=512, dtype=torch.float32, accuracy=4, weak=True),
    "divergence_1024_f32_a4_strong": dict(kind="divergence", n=1024, dtype=torch.float32, accuracy=4, weak=False),
    "curl_1024_f32_a4_strong": dict(kind="curl", n=1024, dtype=torch.float32, accuracy=4, weak=False),
    "laplacian_2048_f64_a8_strong": dict(kind="laplacian", n=2048, dtype=torch.float64, accuracy=8, weak=False),
}
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
names = sys.argv[1:] or ["laplacian_512_f32_a4_weak"]


def timed(fn, steps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for name in names:
    c = CASES[name]
    n = c["n"]
    n0 = n * world if c["weak"] else n
    for mode in ("peer", "nccl"):
        sl = fd.Slab(global_n0=n0, rank=rank, world=world, group=dist.group.WORLD)
        op = fd.SlabOperator(sl, c["kind"], spatial=(n0, n, n), dtype=c["dtype"], accuracy=c["accuracy"], step_size=1.0 / (n - 1), device=dev, n_buffers=1, halo=mode)
        op.fill_random(0, seed=rank)
        eager = timed(lambda: op.step(0))
        marks = {k: [] for k in ("halo", "shell", "interior", "done")}
        for _ in range(20):
            tl = {}
            op.step(0, timeline=tl)
            torch.cuda.synchronize()
            for k in marks:
                if k in tl:
                    marks[k].append(tl["start"].elapsed_time(tl[k]) * 1e3)
        alone = timed(lambda: op.step_alone(0))
        graph = None
        if op.mode == "peer":
            op.capture(0)
            graph = timed(lambda: op.replay(0))
            op.release_graphs()
        rec = dict(case=name, n_gpus=world, rank=rank, halo_transport=op.mode, requested=mode, eager_step_us=round(eager * 1e3, 1),
                   graph_step_us=None if graph is None else round(graph * 1e3, 1), slab_alone_us=round(alone * 1e3, 1),
                   us_after_step_start={k: round(statistics.median(v), 1) for k, v in marks.items() if v},
                   halo_bytes_per_step=int(op.halo_bytes_per_step), peer_error=getattr(op, "peer_error", None))
        out = [None] * world
        dist.all_gather_object(out, rec)
        if rank == 0:
            for r in out:
                print(json.dumps(r), flush=True)
        del op
        torch.cuda.empty_cache()
        dist.barrier()
dist.destroy_process_group()
