"""Profiling driver: one GPR loss+grad evaluation (or only the Cholesky) at a given N.  Dev tool for ncu."""
import sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
what = sys.argv[2] if len(sys.argv) > 2 else "potrf"
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1234)
X = torch.rand(n, 8, generator=g, dtype=torch.float64).to(dev)
ell = torch.ones(8, dtype=torch.float64, device=dev); s2 = torch.ones(1, dtype=torch.float64, device=dev)
noise = torch.full((1,), 0.01, dtype=torch.float64, device=dev)
buf, ld = nv._aligned_empty(n, n, dev)
nv.kern_fwd(0, X, None, ell, s2, noise=noise, lower=True, out=buf, ldk=ld)
dinv, info = nv.potrf_(buf, ld)
if what == "all":
    y = torch.randn(n, 1, dtype=torch.float64, device=dev)
    nv.trsv_(buf, dinv, y, False)
    nv.trsv_(buf, dinv, y, True)
if what in ("potri", "all"):
    kd = nv.potri_(buf, ld, dinv)
if what == "all":
    a = torch.randn(n, 1, dtype=torch.float64, device=dev)
    nv.gpr_grad(0, X, ell, s2, buf, ld, kd, a)
torch.cuda.synchronize()
print("info", info.item())
if what == "all":
    # composite kernel (Linear + Rbf + Constant, BASELINE config #1's kernel) in one pass at the same N
    v = torch.ones(8, dtype=torch.float64, device=dev); c = torch.ones(1, dtype=torch.float64, device=dev)
    out, ldo = nv._aligned_empty(n, n, dev)
    nv.kern_sop_fwd([[(4, v, None)], [(0, ell, s2)], [(6, None, c)]], X, None, noise=noise, lower=True, out=out, ldk=ldo)
    torch.cuda.synchronize()
