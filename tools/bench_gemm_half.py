"""A/B of the 128x128 (1 CTA/SM) and 128x64-half (2 CTAs/SM) GEMM configurations on the shapes of the Cholesky / inverse:
GPB_GEMM_HALF=<min tiles> enables the half configuration, GPB_GEMM_STAGGER=<clocks> overrides the first-wave stagger."""
import os, sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
dev = torch.device("cuda:0")
def timeit(f, n=4):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
out = []
for (m, k) in [(16384, 2048), (28672, 2048), (8192, 2048)]:
    A = torch.randn(m, k, dtype=torch.float64, device=dev); C = torch.randn(m, m, dtype=torch.float64, device=dev)
    C0 = C.clone()
    ms = timeit(lambda: nv.gemm(nv.GEMM_NT, A, A, alpha=-1.0, beta=1.0, C=C, lower_only=True))
    tiles = (m // 128) * (m // 128 + 1) // 2
    out.append("syrk m=%d k=%d %.3f ms %.2f TF" % (m, k, ms, tiles * 128 * 128 * k * 2 / ms / 1e9))
    del A, C, C0
for (m, n, k) in [(131072, 1024, 1024), (28672, 2048, 2048)]:
    A = torch.randn(m, k, dtype=torch.float64, device=dev); B = torch.randn(k, n, dtype=torch.float64, device=dev)
    ms = timeit(lambda: nv.gemm(nv.GEMM_NN, A, B))
    out.append("NN %dx%dx%d %.3f ms %.2f TF" % (m, n, k, ms, 2.0 * m * n * k / ms / 1e9))
# correctness spot check against the 128x128 path is covered by tests; here only a checksum
A = torch.randn(4096, 512, dtype=torch.float64, device=dev, generator=torch.Generator(device="cuda").manual_seed(1))
R = nv.gemm(nv.GEMM_NT, A, A)
out.append("err %.2e" % float((R - A @ A.t()).abs().max() / (A @ A.t()).abs().max()))
print("half=%s stagger=%s | " % (os.environ.get("GPB_GEMM_HALF", "-"), os.environ.get("GPB_GEMM_STAGGER", "auto")) + " | ".join(out), flush=True)
