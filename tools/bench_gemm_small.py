"""M x M products of the sparse models (M = 1024, 2048) in the three operand modes; GPB_GEMM_SMALL_TILES selects up to
how many 128x128 tiles a launch uses 64x64 tiles instead.  Dev tool."""
import os, sys, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
dev = torch.device("cuda:0")
def timeit(f, n=5):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
out = []
for m in (1024, 1536, 2048, 3072):
    A = torch.randn(m, m, dtype=torch.float64, device=dev); B = torch.randn(m, m, dtype=torch.float64, device=dev)
    for name, mode in (("NT", nv.GEMM_NT), ("TN", nv.GEMM_TN), ("NN", nv.GEMM_NN)):
        ms = timeit(lambda: nv.gemm(mode, A, B))
        out.append("%s m=%d %.3f ms %.1f TF" % (name, m, ms, 2.0 * m ** 3 / ms / 1e9))
    ms = timeit(lambda: nv.gemm(nv.GEMM_NT, A, A, lower_only=True))
    out.append("SYRK m=%d %.3f ms" % (m, ms))
print("small_tiles=%s | " % os.environ.get("GPB_GEMM_SMALL_TILES", "default(64)") + " | ".join(out), flush=True)
