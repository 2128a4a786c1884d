"""One gpb_potrf_lower at N (after a warm-up factorisation) for an ncu launch list.  Dev tool."""
import sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1234)
X = torch.rand(n, 8, generator=g, dtype=torch.float64).to(dev)
ell = torch.ones(8, dtype=torch.float64, device=dev); s2 = torch.ones(1, dtype=torch.float64, device=dev)
noise = torch.full((1,), 0.01, dtype=torch.float64, device=dev)
buf, ld = nv._aligned_empty(n, n, dev)
nv.kern_fwd(0, X, None, ell, s2, noise=noise, lower=True, out=buf, ldk=ld)
dinv, info = nv.potrf_(buf, ld)
torch.cuda.synchronize()
print("info", info.item())
