"""Small grids: fused kernel vs composition of axis kernels (CUDA-graph replay = device time per call)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
def dev_time(call, N=300):
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        call()
        with torch.cuda.graph(g, stream=s): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(20): g.replay()
    e0.record()
    for _ in range(N): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N * 1e3
for n in (48, 64, 100, 128, 160, 200, 256, 320):
    F = torch.randn(3, n, n, n, device="cuda"); x = F[0].contiguous(); h = 1.0 / (n - 1)
    row = [f"n={n}"]
    for name, call in (("div a6", lambda: fdx.divergence(F, step_size=h, keepdims=False, accuracy=6)), ("lap a4", lambda: fdx.laplacian(x, step_size=h, accuracy=4)),
                       ("grad a4", lambda: fdx.gradient(x, step_size=h, accuracy=4)), ("curl a4", lambda: fdx.curl(F, step_size=h, accuracy=4))):
        os.environ["FDX_FUSED_DISABLE"] = "0"; tf = dev_time(call)
        os.environ["FDX_FUSED_DISABLE"] = "1"; tc = dev_time(call)
        row.append(f"{name}: fused {tf:.1f} us, composed {tc:.1f} us")
    print(" | ".join(row), flush=True)
os.environ["FDX_FUSED_DISABLE"] = "0"
for n in (64, 100, 160, 256):
    F = torch.randn(3, n, n, n, device="cuda"); x = F[0].contiguous(); h = 1.0 / (n - 1)
    row = [f"n={n} uniform chunk sweep"]
    for ch in (4, 6, 8, 12, 16, 24, 32):
        os.environ["FDX_FUSED_CHUNK"] = str(ch)
        row.append(f"{ch}: div {dev_time(lambda: fdx.divergence(F, step_size=h, keepdims=False, accuracy=6)):.1f} lap {dev_time(lambda: fdx.laplacian(x, step_size=h, accuracy=4)):.1f}")
    print(" | ".join(row), flush=True)
