import sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
dev = torch.device("cuda:0"); torch.manual_seed(0)
nbad = 0
for rep in range(6):
    for (m, n, k) in [(2052, 2052, 2048), (2048, 2048, 2048), (4096, 2048, 4096)]:
        A = torch.randn(m, k, dtype=torch.float64, device=dev); B = torch.randn(n, k, dtype=torch.float64, device=dev)
        C0 = torch.randn(m, n, dtype=torch.float64, device=dev)
        buf, ld = nv._aligned_empty(m, n, dev); C = buf[:, :n]; C.copy_(C0)
        nv.gemm(nv.GEMM_NT, A, B, alpha=-1.0, beta=1.0, C=C)
        ref = C0 - A @ B.t()
        bad = (C - ref).abs() > 1e-9
        if bad.any():
            nbad += 1
            idx = bad.nonzero()
            print(f"rep {rep} m={m} n={n} k={k}: BAD count {bad.sum().item()} rows {idx[:,0].min().item()}..{idx[:,0].max().item()} cols {idx[:,1].min().item()}..{idx[:,1].max().item()}")
print("total bad", nbad)
