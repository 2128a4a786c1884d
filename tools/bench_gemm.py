"""Throughput of the DMMA GEMM engine on shapes the Cholesky uses.  Dev tool."""
import sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
dev = torch.device("cuda:0")
def timeit(f, n=3):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
big = torch.randn(30720, 30720, dtype=torch.float64, device=dev)
print("shape                                  ms      TFLOP/s")
for (m, k) in [(28672, 2048), (16384, 2048), (8192, 2048), (4096, 2048), (16384, 16384), (16384, 8192), (16384, 4096), (16384, 1024), (16384, 512), (28672, 1024), (28672, 4096)]:
    A = big[:m, :k]; C = big[:m, m - 2048 if False else 0:][:, :m] if False else torch.empty(m, m, dtype=torch.float64, device=dev)
    ms = timeit(lambda: nv.gemm(nv.GEMM_NT, A, A, alpha=-1.0, beta=1.0, C=C, lower_only=True))
    tiles = (m // 128) * (m // 128 + 1) // 2
    print(f"syrk  m=n={m:6d} k={k:6d} tiles={tiles:6d}  {ms:9.3f} {tiles*128*128*k*2/ms/1e9:8.2f}")
    del C
for (m, n, k) in [(28672, 2048, 2048), (16384, 2048, 2048), (8192, 2048, 2048), (28672, 128, 128), (28672, 1024, 1024), (28672, 256, 128), (28672, 512, 256), (16384, 128, 128), (2048, 2048, 2048), (1024, 1024, 1024), (2048, 128, 128)]:
    A = big[:m, :k]; B = big[:n, k:2 * k]; C = torch.empty(m, n, dtype=torch.float64, device=dev)
    ms = timeit(lambda: nv.gemm(nv.GEMM_NT, A, B, alpha=-1.0, beta=1.0, C=C))
    print(f"gemm  m={m:6d} n={n:6d} k={k:6d} tiles={(m//128)*max(n//128,1):6d} {ms:9.3f} {2.0*m*n*k/ms/1e9:8.2f}")
