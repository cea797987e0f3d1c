"""Multi-GPU slab operator == single-GPU operator on the gathered grid (run under torchrun, N ranks):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import distributed as fd

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
for kind, dtype, acc, n0, n1, n2 in (("laplacian", torch.float32, 4, 64 * world, 80, 192), ("divergence", torch.float64, 2, 40 * world, 48, 96),
                                     ("curl", torch.float32, 4, 48 * world, 64, 128), ("gradient", torch.float32, 6, 56 * world, 48, 200)):
    sl = fd.Slab(global_n0=n0, rank=rank, world=world, group=dist.group.WORLD)
    h = (0.1, 0.2, 0.3)
    op = fd.SlabOperator(sl, kind, spatial=(n0, n1, n2), dtype=dtype, accuracy=acc, step_size=h, device=dev, n_buffers=2)
    for b in range(2):
        op.fill_random(b, seed=7 * rank + b + 1)
    for rep in range(3):
        for b in range(2):
            outs = op.step(b)
    torch.cuda.synchronize()
    for b in range(2):
        outs = op.step(b)
        nin = len(op.inputs[b])
        full_in = []
        for f in range(nin):
            parts = [torch.empty_like(op.owned_view(b, f)) for _ in range(world)]
            dist.all_gather(parts, op.owned_view(b, f).contiguous())
            full_in.append(torch.cat(parts, 0))
        x = full_in[0] if nin == 1 else torch.stack(full_in, 0)
        if kind == "divergence":
            ref = fdx.divergence(x, accuracy=acc, step_size=h, keepdims=False)[None]
        elif kind == "laplacian":
            ref = fdx.laplacian(x, accuracy=acc, step_size=h)[None]
        else:
            ref = getattr(fdx, kind)(x, accuracy=acc, step_size=h)
        for o, t in enumerate(outs):
            mine = ref[o][sl.z0 : sl.z1]
            if not torch.equal(t, mine):
                ok = False
                print(f"rank {rank}: MISMATCH {kind} buffer {b} output {o}: max |d| = {(t - mine).abs().max().item():.3e}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("dist_check", "OK (bit-identical to the single-GPU operator)" if int(flag.item()) else "FAILED", flush=True)
dist.destroy_process_group()
