"""One SYRK trailing update (the dominant launch shape of the Cholesky) for `ncu --set full`.  Dev tool."""
import sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
dev = torch.device("cuda:0")
m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
A = torch.randn(m, k, dtype=torch.float64, device=dev)
C = torch.randn(m, m, dtype=torch.float64, device=dev)
for _ in range(2):
    nv.gemm(nv.GEMM_NT, A, A, alpha=-1.0, beta=1.0, C=C, lower_only=True)
torch.cuda.synchronize()
