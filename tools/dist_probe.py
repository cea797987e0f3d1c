"""CPU vs GPU time of SlabOperator.step at N ranks (torchrun): python -m torch.distributed.run ... tools/dist_probe.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finitediffx_b200 import distributed as fd
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 512
sl = fd.Slab(global_n0=n * world, rank=rank, world=world, group=dist.group.WORLD)
op = fd.SlabOperator(sl, "laplacian", spatial=(n, n, n), dtype=torch.float32, accuracy=4, step_size=1.0 / (n - 1), device=dev, n_buffers=3)
for b in range(3): op.fill_random(b, seed=rank * 10 + b)
for i in range(10): op.step(i % 3)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
K = 200
t0 = time.perf_counter()
for i in range(K): op.step(i % 3)
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
if rank == 0: print(f"world {world}: CPU enqueue {t_cpu/K*1e6:.1f} us/step, total {t_all/K*1e6:.1f} us/step")
graph_ok = os.environ.get("PROBE_GRAPH", "0") == "1"  # (the process-group teardown hangs after an NCCL capture on this stack)
if graph_ok:
    try:
        graphs = []
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for b in range(3):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    op.step(b)
                graphs.append(g)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        for i in range(10): graphs[i % 3].replay()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K): graphs[i % 3].replay()
        t_cpu = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        if rank == 0: print(f"world {world} graphs: CPU enqueue {t_cpu/K*1e6:.1f} us/step, total {t_all/K*1e6:.1f} us/step")
    except Exception as e:
        if rank == 0: print("graph capture failed:", repr(e)[:300])
dist.destroy_process_group()
