import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import finitediffx_b200 as fdx
from finitediffx_b200 import plan as P, _cabi
from oracle import fdx_oracle as O
from tests import _golden
from tests._parity import parity_error, scale_for
c=[c for c in _golden.load("x32") if c["id"]==181][0]
print(c["op"], c["kwargs"], c["x"].shape, c["x"].dtype)
P.set_coefficient_source(lambda offs, d: O.coeffs_reference(offs, d, x64=False).astype(np.float64))
x=torch.from_numpy(c["x"]).cuda()
for dis in ("1","0"):
    os.environ["FDX_FUSED_DISABLE"]=dis
    out=fdx.laplacian(x, **c["kwargs"]).cpu().numpy()
    s=scale_for("laplacian", c["x"], c["kwargs"]); e=np.abs(out.astype(np.float64)-c["out"])/s
    idx=np.unravel_index(np.argmax(e), e.shape)
    print("disable",dis,_cabi.last_kernel(),"max err",e.max(),"at",idx, "count>1e-5", (e>1e-5).sum(), "of", e.size)
    bad=np.argwhere(e>1e-5)
    print("  bad z range",bad[:,0].min() if len(bad) else None, bad[:,0].max() if len(bad) else None, "y",bad[:,1].min() if len(bad) else None,bad[:,1].max() if len(bad) else None,"x",bad[:,2].min() if len(bad) else None,bad[:,2].max() if len(bad) else None)
ref=O.laplacian(c["x"], coeff_mode="ref32", **c["kwargs"])
print("oracle ref32 vs golden equal:", np.array_equal(ref, c["out"]))
