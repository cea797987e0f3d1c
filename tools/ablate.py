import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
n=512; xs=[torch.randn(n,n,n,device="cuda") for _ in range(3)]; h=1.0/(n-1)
cfgs=sys.argv[1].split(",") if len(sys.argv)>1 else ["11"]
abls=[int(a) for a in sys.argv[2].split(",")] if len(sys.argv)>2 else [0,31,32,63]
for cfg in cfgs:
  os.environ["FDX_FUSED_CFG"]=cfg; os.environ["FDX_FUSED_CHUNK"]="32"
  for ab in abls:
    os.environ["FDX_ABLATE"]=str(ab)
    for i in range(3): fdx.laplacian(xs[i%3],accuracy=4,step_size=h)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): fdx.laplacian(xs[i%3],accuracy=4,step_size=h)
    e1.record(); torch.cuda.synchronize()
    print(f"cfg {cfg} ablate {ab:2d} {e0.elapsed_time(e1)/20:.4f} ms", flush=True)
