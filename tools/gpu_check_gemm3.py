import sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = n = 2052; k = 2048
for rep in range(8):
    A = torch.randn(m, k, dtype=torch.float64, device=dev); B = torch.randn(n, k, dtype=torch.float64, device=dev)
    C0 = torch.randn(m, n, dtype=torch.float64, device=dev)
    buf, ld = nv._aligned_empty(m, n, dev); C = buf[:, :n]; C.copy_(C0)
    nv.gemm(nv.GEMM_NT, A, B, alpha=-1.0, beta=1.0, C=C)
    P = A @ B.t()
    ref = C0 - P
    bad = (C - ref).abs() > 1e-9
    if bad.any():
        idx = bad.nonzero()
        rows = idx[:, 0].unique().tolist(); cols = idx[:, 1].unique()
        print(f"rep {rep}: count {bad.sum().item()} rows {rows} cols {cols.min().item()}..{cols.max().item()} ({cols.numel()} cols)")
        for (i, j) in idx[:6].tolist():
            print(f"   ({i},{j}) got {C[i,j].item():+.6f} ref {ref[i,j].item():+.6f} C0 {C0[i,j].item():+.6f} -P {-P[i,j].item():+.6f} got-ref {C[i,j].item()-ref[i,j].item():+.6f}")
        # does the wrong value match ref of another element?
        i, j = idx[0].tolist()
        d = C[i, j] - ref[i, j]
        near = ((C0 - d).abs() < 1e-9).nonzero()
        print("   delta equals C0 at", near[:5].tolist(), " -delta equals C0 at", ((C0 + d).abs() < 1e-9).nonzero()[:5].tolist())
