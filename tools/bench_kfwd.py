"""Covariance forward pass alone (gpb_kern_fwd): GPR fill (N x N lower, D = 8) and one sparse panel (rows x M, D = 16).
Dev tool; GPB_KFWD_CFG selects the D <= 16 kernel configuration under test.  Prints a device-side integer checksum of
every output so that configurations can be compared bit for bit."""
import os
import sys
import torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
g = torch.Generator().manual_seed(1234)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


def checksum(t):
    return int(t.view(torch.int64).sum().item())


print("GPB_KFWD_CFG =", os.environ.get("GPB_KFWD_CFG", "(default)"))
X = torch.rand(n, 8, generator=g, dtype=torch.float64).to(dev)
ell = torch.full((8,), 0.7, dtype=torch.float64, device=dev)
s2 = torch.full((1,), 1.3, dtype=torch.float64, device=dev)
noise = torch.full((1,), 0.01, dtype=torch.float64, device=dev)
buf, ld = nv._aligned_empty(n, n, dev)
buf.zero_()
for kind, name in ((0, "rbf"), (1, "exp"), (2, "matern32"), (3, "matern52")):
    ms = timed(lambda: nv.kern_fwd(kind, X, None, ell, s2, noise=noise, lower=True, out=buf, ldk=ld))
    print("gpr fill  %-9s N=%d D=8 lower : %.3f ms   checksum %d" % (name, n, ms, checksum(buf)))
del buf
rows, m = 131072, 1024
X16 = torch.rand(rows, 16, generator=g, dtype=torch.float64).to(dev)
Z = torch.rand(m, 16, generator=g, dtype=torch.float64).to(dev)
ell16 = torch.full((16,), 1.5, dtype=torch.float64, device=dev)
out, ldo = nv._aligned_empty(rows, m, dev)
for kind, name in ((0, "rbf"), (3, "matern52")):
    ms = timed(lambda: nv.kern_fwd(kind, X16, Z, ell16, s2, out=out, ldk=ldo))
    print("panel     %-9s %d x %d D=16     : %.3f ms   checksum %d" % (name, rows, m, ms, checksum(out)))
