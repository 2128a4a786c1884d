"""Launch shapes of the sparse models for `ncu --set full`: the split-K TN Gram product of a VFE panel
(Phi += Kfu^T Kfu, 131072 x 1024), the NN panel product G = Kfu R and the covariance backward on that panel.  Dev tool."""
import sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
dev = torch.device("cuda:0")
rows, m, d = 131072, 1024, 16
g = torch.Generator().manual_seed(0)
X = torch.rand(rows, d, generator=g, dtype=torch.float64).to(dev)
Z = torch.rand(m, d, generator=g, dtype=torch.float64).to(dev)
ell = torch.ones(d, dtype=torch.float64, device=dev); s2 = torch.ones(1, dtype=torch.float64, device=dev)
P = nv.kern_fwd(0, X, Z, ell, s2)
AA3 = torch.zeros((16, m, m), dtype=torch.float64, device=dev)
R = torch.randn(m, m, dtype=torch.float64, device=dev)
for _ in range(2):
    nv.gemm_splitk(nv.GEMM_TN, P, P, 8192, AA3, beta=1.0, lower_only=True)
    G = nv.gemm(nv.GEMM_NN, P, R)
    nv.kern_bwd(0, X, Z, ell, s2, G, True)
torch.cuda.synchronize()
