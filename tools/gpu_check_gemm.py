import sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
dev = torch.device("cuda:0"); torch.manual_seed(0)
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
for (m, n, k) in [(2052, 2052, 2048), (1028, 1024, 512), (4100, 4100, 128), (1156, 1280, 256), (2052, 1024, 1024), (1028, 2052, 64), (2048, 2048, 256)]:
    A = torch.randn(m, k, dtype=torch.float64, device=dev); B = torch.randn(n, k, dtype=torch.float64, device=dev)
    C0 = torch.randn(m, n, dtype=torch.float64, device=dev)
    for beta in (0.0, 1.0):
        buf, ld = nv._aligned_empty(m, n, dev); C = buf[:, :n]; C.copy_(C0)
        nv.gemm(nv.GEMM_NT, A, B, alpha=-1.0, beta=beta, C=C)
        ref = beta * C0 - A @ B.t()
        err = rel(C, ref)
        bad = (C - ref).abs() > 1e-9
        print(f"NT m={m} n={n} k={k} beta={beta}: rel {err:.2e}", "" if err < 1e-12 else f"BAD rows {bad.nonzero()[:,0].unique()[:8].tolist()} cols {bad.nonzero()[:,1].unique()[:8].tolist()} count {bad.sum().item()}")
        At = A.t().contiguous(); Bt = B.t().contiguous()
        buf, ld = nv._aligned_empty(m, n, dev); C = buf[:, :n]; C.copy_(C0)
        nv.gemm(nv.GEMM_TN, At, Bt, alpha=-1.0, beta=beta, C=C)
        err = rel(C, ref); print(f"TN ... rel {err:.2e}", "" if err < 1e-12 else "BAD")
        buf, ld = nv._aligned_empty(m, n, dev); C = buf[:, :n]; C.copy_(C0)
        nv.gemm(nv.GEMM_NN, A, Bt, alpha=-1.0, beta=beta, C=C)
        err = rel(C, ref); print(f"NN ... rel {err:.2e}", "" if err < 1e-12 else "BAD")
    if m == n:
        buf, ld = nv._aligned_empty(m, n, dev); C = buf[:, :n]; C.copy_(C0)
        nv.gemm(nv.GEMM_NT, A, A, alpha=-1.0, beta=1.0, C=C, lower_only=True)
        ref = C0 - A @ A.t()
        err = rel(torch.tril(C), torch.tril(ref)); print(f"SYRK lower m={m} k={k}: rel {err:.2e}", "" if err < 1e-12 else "BAD")
