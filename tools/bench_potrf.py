"""Time gpb_potrf_lower at a given N (best of 3, CUDA events) -- run once per GPB_LA_* setting (read at first use).

    GPB_LA_PANEL=2048 GPB_LA_FIRST=1024 GPB_LA_TAIL=4096 python tools/bench_potrf.py 32768
"""
import os, sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1234)
X = torch.rand(n, 8, generator=g, dtype=torch.float64).to(dev)
ell = torch.ones(8, dtype=torch.float64, device=dev); s2 = torch.ones(1, dtype=torch.float64, device=dev)
noise = torch.full((1,), 0.01, dtype=torch.float64, device=dev)
buf, ld = nv._aligned_empty(n, n, dev)
best = 1e30
for it in range(4):
    nv.kern_fwd(0, X, None, ell, s2, noise=noise, lower=True, out=buf, ldk=ld)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dinv, info = nv.potrf_(buf, ld); e1.record(); torch.cuda.synchronize()
    if it: best = min(best, e0.elapsed_time(e1))
ld_sum = nv.logdet_sumsq(buf)[0].item()
print("n=%d panel=%s first=%s tail=%s: potrf %.2f ms = %.2f TFLOP/s  info=%d logdet=%.10f" % (
    n, os.environ.get("GPB_LA_PANEL", "-"), os.environ.get("GPB_LA_FIRST", "-"), os.environ.get("GPB_LA_TAIL", "-"),
    best, n ** 3 / 3 / best / 1e9, info.item(), ld_sum), flush=True)
