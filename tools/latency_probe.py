"""Config 1 (README example, 100^3 divergence accuracy 6): where does a call's time go?"""
import os, sys, time, cProfile, pstats, io
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
F = torch.randn(3, 100, 100, 100, device="cuda"); h = 1.0 / 99
call = lambda: fdx.divergence(F, step_size=(h, h, h), keepdims=False, accuracy=6, method="central")
for _ in range(50): call()
torch.cuda.synchronize()
N = 2000
t0 = time.perf_counter()
for _ in range(N): call()
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue {t_cpu/N*1e6:.1f} us/call, with GPU drain {t_all/N*1e6:.1f} us/call")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    call()
    with torch.cuda.graph(g, stream=s):
        out = call()
torch.cuda.synchronize()
e0.record()
for _ in range(N): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"CUDA-graph replay {e0.elapsed_time(e1)/N*1e3:.1f} us/call (kernel + launch only)")
pr = cProfile.Profile(); pr.enable()
for _ in range(N): call()
pr.disable(); torch.cuda.synchronize()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(14); print(st.getvalue()[:3000])
