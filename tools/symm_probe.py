"""Does torch symmetric memory (CUDA backend: peer-mapped buffers over NVLink) work on this box?  Run under torchrun.
Allocates a buffer per rank, rendezvous, reads the neighbour's buffer through the mapped pointer with a plain copy and
times a peer D2D copy of a halo-sized block and the signal-pad barrier."""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 16 * 512 * 512
t = symm.empty(n, dtype=torch.float32, device=dev)
t.fill_(float(rank + 1))
hdl = symm.rendezvous(t, dist.group.WORLD)
print(f"rank {rank}: rendezvous ok, world {hdl.world_size}, multicast {getattr(hdl, 'has_multicast_support', None)}", flush=True)
hdl.barrier(channel=0)
peer = (rank + 1) % world
pbuf = hdl.get_buffer(peer, (n,), torch.float32)
local = torch.empty(n, device=dev)
local.copy_(pbuf)
torch.cuda.synchronize()
assert float(local[0]) == float(peer + 1) and float(local[-1]) == float(peer + 1), (float(local[0]), peer + 1)
print(f"rank {rank}: peer read ok (ptr {pbuf.data_ptr():#x})", flush=True)
hdl.barrier(channel=0)
for nbytes in (3 << 20, 16 << 20):
    k = nbytes // 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        local[:k].copy_(pbuf[:k])
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        local[:k].copy_(pbuf[:k])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"rank {rank}: peer copy {nbytes >> 20} MiB: {us:.1f} us, {nbytes / us / 1e3:.0f} GB/s", flush=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    hdl.barrier(channel=0)
e0.record()
for _ in range(50):
    hdl.barrier(channel=0)
e1.record()
torch.cuda.synchronize()
print(f"rank {rank}: symm barrier {e0.elapsed_time(e1) / 50 * 1e3:.1f} us", flush=True)
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank}: done", flush=True)
