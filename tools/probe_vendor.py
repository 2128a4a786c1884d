"""Vendor-library bar on the B200 (cuBLAS dgemm / cuSOLVER potrf, potri / trsm). Not product code."""
import torch, time, json, sys
dev = torch.device("cuda:0")
def t(f, n=3):
    f(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {}
for n in (4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = t(lambda: torch.matmul(a, b)); out[f"dgemm_nn_{n}"] = 2 * n**3 / ms / 1e9
    ms = t(lambda: torch.matmul(a, b.t())); out[f"dgemm_nt_{n}"] = 2 * n**3 / ms / 1e9
    print(n, out, flush=True)
    del a, b
for n in (8192, 16384, 32768):
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(n, 8, dtype=torch.float64, device=dev, generator=g)
    d = torch.cdist(x, x) ** 2
    k = torch.exp(-0.5 * d); del d
    k.diagonal().add_(0.01)
    ms = t(lambda: torch.linalg.cholesky(k), 2); out[f"potrf_{n}_tflops"] = n**3 / 3 / ms / 1e9; out[f"potrf_{n}_ms"] = ms
    L = torch.linalg.cholesky(k); del k
    ms = t(lambda: torch.cholesky_inverse(L), 2); out[f"potri_{n}_tflops"] = 2 * n**3 / 3 / ms / 1e9; out[f"potri_{n}_ms"] = ms
    y = torch.randn(n, 1, dtype=torch.float64, device=dev)
    ms = t(lambda: torch.linalg.solve_triangular(L, y, upper=False)); out[f"trsv_{n}_ms"] = ms; out[f"trsv_{n}_gbs"] = 4 * n * n / ms / 1e6
    if n <= 16384:
        B = torch.randn(n, n, dtype=torch.float64, device=dev)
        ms = t(lambda: torch.linalg.solve_triangular(L, B, upper=False), 2); out[f"trsm_{n}_tflops"] = n**3 / ms / 1e9
        del B
    print(n, {k_: v for k_, v in out.items() if str(n) in k_}, flush=True)
    del L, y
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/probe_vendor.json", "w"), indent=1)
