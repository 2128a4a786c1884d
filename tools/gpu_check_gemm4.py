import sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = n = 2052; k = 2048
for rep in range(10):
    A = torch.randn(m, k, dtype=torch.float64, device=dev); B = torch.randn(n, k, dtype=torch.float64, device=dev)
    C0 = torch.randn(m, n, dtype=torch.float64, device=dev)
    A1, B1 = A.clone(), B.clone()
    P = A @ B.t()
    ref = C0 - P
    buf, ld = nv._aligned_empty(m, n, dev); C = buf[:, :n]; C.copy_(C0)
    torch.cuda.synchronize()
    nv.gemm(nv.GEMM_NT, A, B, alpha=-1.0, beta=1.0, C=C)
    torch.cuda.synchronize()
    bad = (C - ref).abs() > 1e-9
    print(f"rep {rep}: bad {bad.sum().item()} A changed {(A != A1).sum().item()} B changed {(B != B1).sum().item()}", end=" ")
    if bad.any():
        idx = bad.nonzero(); print("rows", idx[:,0].unique().tolist()[:12], "cols", idx[:,1].min().item(), idx[:,1].max().item(), "tiles_n", sorted(set((idx[:,1]//128).tolist())))
    else: print()
