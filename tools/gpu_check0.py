"""First-light check of every native primitive against torch CUDA (cuBLAS/cuSOLVER).  Dev tool, not a test."""
import sys, time, torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv
dev = torch.device("cuda:0")
torch.manual_seed(0)
def rel(a, b): return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()
def timeit(f, n=3):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ok = True
def report(name, err, tol):
    global ok
    flag = "OK " if err <= tol else "BAD"
    if not err <= tol: ok = False
    print(f"{flag} {name}: rel err {err:.3e}", flush=True)

# ---- gemm
for (m, n, k) in [(128, 128, 16), (256, 128, 128), (300, 200, 100), (1024, 768, 512), (77, 130, 1000)]:
    A = torch.randn(m, k, dtype=torch.float64, device=dev); B = torch.randn(n, k, dtype=torch.float64, device=dev)
    report(f"gemm NT {m}x{n}x{k}", rel(nv.gemm(nv.GEMM_NT, A, B), A @ B.t()), 1e-13)
    At = A.t().contiguous(); Bt = B.t().contiguous()
    report(f"gemm TN {m}x{n}x{k}", rel(nv.gemm(nv.GEMM_TN, At, Bt), A @ B.t()), 1e-13)
    report(f"gemm NN {m}x{n}x{k}", rel(nv.gemm(nv.GEMM_NN, A, Bt), A @ B.t()), 1e-13)
A = torch.randn(512, 256, dtype=torch.float64, device=dev)
C0 = torch.randn(512, 512, dtype=torch.float64, device=dev); C = C0.clone()
nv.gemm(nv.GEMM_NT, A, A, alpha=-1.0, beta=1.0, C=C, lower_only=True)
ref = C0 - A @ A.t()
report("syrk lower", rel(torch.tril(C), torch.tril(ref)), 1e-13)

# ---- kern fwd
def ref_K(kind, X, X2, ell, s2):
    Xs = X / ell; X2s = (X if X2 is None else X2) / ell
    r2 = (Xs**2).sum(1, keepdim=True) + (X2s**2).sum(1, keepdim=True).t() - 2 * Xs @ X2s.t()
    r2 = r2.clamp_min(0)
    if X2 is None: r2 = r2 * (1.0 - torch.eye(X.shape[0], dtype=r2.dtype, device=r2.device))   # exact zero self-distance
    if kind == 0: return s2 * torch.exp(-0.5 * r2)
    r = torch.sqrt(r2.clamp_min(1e-40))
    if kind == 1: return s2 * torch.exp(-r)
    if kind == 2: return s2 * (1 + 3**0.5 * r) * torch.exp(-3**0.5 * r)
    return s2 * (1 + 5**0.5 * r + 5.0 / 3 * r * r) * torch.exp(-5**0.5 * r)
for D in (1, 3, 8, 16, 33):
    X = torch.rand(333, D, dtype=torch.float64, device=dev); X2 = torch.rand(257, D, dtype=torch.float64, device=dev)
    ell = 0.3 + torch.rand(D, dtype=torch.float64, device=dev); s2 = torch.tensor([1.7], dtype=torch.float64, device=dev)
    for kind in range(4):
        report(f"kern_fwd kind={kind} D={D} X,X2", rel(nv.kern_fwd(kind, X, X2, ell, s2), ref_K(kind, X, X2, ell, s2)), 1e-12)
        report(f"kern_fwd kind={kind} D={D} sym iso", rel(nv.kern_fwd(kind, X, None, ell[:1], s2), ref_K(kind, X, None, ell[:1], s2)), 1e-12)
    v = ell
    report(f"kern_fwd linear D={D}", rel(nv.kern_fwd(4, X, X2, v, None), (X * v) @ X2.t()), 1e-13)
noise = torch.tensor([0.01], dtype=torch.float64, device=dev)
X = torch.rand(700, 8, dtype=torch.float64, device=dev); ell = torch.ones(8, dtype=torch.float64, device=dev); s2 = torch.ones(1, dtype=torch.float64, device=dev)
Kl = nv.kern_fwd(0, X, None, ell, s2, noise=noise, lower=True)
Kr = ref_K(0, X, None, ell, s2) + 0.01 * torch.eye(700, dtype=torch.float64, device=dev)
report("kern_fwd lower+noise", rel(torch.tril(Kl), torch.tril(Kr)), 1e-13)

# ---- potrf / trsv / logdet / potri
for n in (5, 100, 128, 129, 300, 1000, 2048, 4100, 6400, 8300):
    X = torch.rand(n, 8, dtype=torch.float64, device=dev)
    K = ref_K(0, X, None, torch.ones(8, dtype=torch.float64, device=dev), 1.0) + 0.01 * torch.eye(n, dtype=torch.float64, device=dev)
    Lref = torch.linalg.cholesky(K)
    buf, ld = nv.sym_buffer_from(K)
    dinv, info = nv.potrf_(buf, ld)
    L = torch.tril(buf[:, :n])
    report(f"potrf n={n} (info={info.item()})", rel(L, Lref), 1e-11)
    y = torch.randn(n, 3, dtype=torch.float64, device=dev)
    x1 = nv.trsv_(buf, dinv, y.clone(), False)
    report(f"trsv fwd n={n}", rel(x1, torch.linalg.solve_triangular(Lref, y, upper=False)), 1e-10)
    x2 = nv.trsv_(buf, dinv, y.clone(), True)
    report(f"trsv bwd n={n}", rel(x2, torch.linalg.solve_triangular(Lref.t(), y, upper=True)), 1e-10)
    ls = nv.logdet_sumsq(buf, x1)
    report(f"logdet n={n}", abs(ls[0].item() - Lref.diagonal().log().sum().item()) / abs(Lref.diagonal().log().sum().item()), 1e-12)
    report(f"sumsq n={n}", abs(ls[1].item() - (x1**2).sum().item()) / (x1**2).sum().item(), 1e-12)
    P = torch.randn(200, n, dtype=torch.float64, device=dev)
    Pb, ldp = nv._aligned_empty(200, n, dev); Pb[:, :n].copy_(P)
    nv.trsm_right_lt_(buf, dinv, Pb, ldp)
    report(f"trsm_right n={n}", rel(Pb[:, :n], torch.linalg.solve_triangular(Lref, P.t(), upper=False).t()), 1e-10)
    a = torch.cholesky_solve(y, Lref)
    kd = nv.potri_(buf, ld, dinv)
    Kinv = nv.potri_assemble(buf, ld, kd)
    Kinv_ref = torch.cholesky_inverse(Lref)
    report(f"potri n={n}", rel(Kinv, Kinv_ref), 1e-9)
    # fused gradient vs autograd
    for kind in (0, 2, 3):
        ellp = (0.5 + torch.rand(8, dtype=torch.float64, device=dev)).requires_grad_(True)
        s2p = torch.tensor([1.3], dtype=torch.float64, device=dev, requires_grad=True)
        W = 0.5 * (3 * Kinv_ref - a @ a.t())
        Kk = ref_K(kind, X, None, ellp, s2p)
        (W * Kk).sum().backward()
        g_ell, g_s2, g_n = nv.gpr_grad(kind, X, ellp.detach(), s2p.detach(), buf, ld, kd, a)
        report(f"gpr_grad kind={kind} n={n} ell", rel(g_ell, ellp.grad), 1e-8)
        report(f"gpr_grad kind={kind} n={n} s2", rel(g_s2, s2p.grad), 1e-8)
        report(f"gpr_grad kind={kind} n={n} noise", abs(g_n.item() - W.diagonal().sum().item()) / abs(W.diagonal().sum().item()), 1e-8)

# ---- kern_bwd dense
for D in (3, 8, 20):
    X = torch.rand(150, D, dtype=torch.float64, device=dev); X2 = torch.rand(333, D, dtype=torch.float64, device=dev)
    G = torch.randn(150, 333, dtype=torch.float64, device=dev)
    for kind in range(4):
        for ard in (True, False):
            ellp = (0.5 + torch.rand(D if ard else 1, dtype=torch.float64, device=dev)).requires_grad_(True)
            s2p = torch.tensor([1.3], dtype=torch.float64, device=dev, requires_grad=True)
            X2p = X2.clone().requires_grad_(True); Xp = X.clone().requires_grad_(True)
            (ref_K(kind, Xp, X2p, ellp, s2p) * G).sum().backward()
            g_ell, g_s2, gX2 = nv.kern_bwd(kind, X, X2, ellp.detach(), s2p.detach(), G, True)
            tol = 1e-9 if kind != 1 else 1e-6
            report(f"kern_bwd kind={kind} D={D} ard={ard} ell", rel(g_ell, ellp.grad), tol)
            report(f"kern_bwd kind={kind} D={D} ard={ard} s2", rel(g_s2, s2p.grad), tol)
            report(f"kern_bwd kind={kind} D={D} ard={ard} gX2", rel(gX2, X2p.grad), tol)
            _, _, gX1 = nv.kern_bwd(kind, X2, X, ellp.detach(), s2p.detach(), G, True, g_transposed=True)
            report(f"kern_bwd kind={kind} D={D} ard={ard} gX1", rel(gX1, Xp.grad), tol)
    vp = (0.5 + torch.rand(D, dtype=torch.float64, device=dev)).requires_grad_(True)
    X2p = X2.clone().requires_grad_(True)
    (((X * vp) @ X2p.t()) * G).sum().backward()
    g_v, _, gX2 = nv.kern_bwd(4, X, X2, vp.detach(), None, G, True)
    report(f"kern_bwd linear D={D} v", rel(g_v, vp.grad), 1e-10)
    report(f"kern_bwd linear D={D} gX2", rel(gX2, X2p.grad), 1e-10)

# ---- timing at scale
if "--big" in sys.argv:
    for n in (8192, 16384, 32768):
        X = torch.rand(n, 8, dtype=torch.float64, device=dev)
        ell = torch.ones(8, dtype=torch.float64, device=dev); s2 = torch.ones(1, dtype=torch.float64, device=dev)
        buf, ld = nv._aligned_empty(n, n, dev)
        ms_k = timeit(lambda: nv.kern_fwd(0, X, None, ell, s2, noise=noise, lower=True, out=buf, ldk=ld), 2)
        t0 = time.time()
        ms_p = timeit(lambda: (nv.kern_fwd(0, X, None, ell, s2, noise=noise, lower=True, out=buf, ldk=ld), nv.potrf_(buf, ld)), 2) - ms_k
        dinv, info = nv.potrf_(buf, ld)   # buffer now holds L of L (garbage but PD-ish?) -> redo properly
        nv.kern_fwd(0, X, None, ell, s2, noise=noise, lower=True, out=buf, ldk=ld); dinv, info = nv.potrf_(buf, ld)
        y = torch.randn(n, 1, dtype=torch.float64, device=dev)
        ms_t = timeit(lambda: nv.trsv_(buf, dinv, y.clone(), False), 2)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); kd = nv.potri_(buf, ld, dinv); e1.record(); torch.cuda.synchronize(); ms_i = e0.elapsed_time(e1)
        a = torch.randn(n, 1, dtype=torch.float64, device=dev)
        ms_g = timeit(lambda: nv.gpr_grad(0, X, ell, s2, buf, ld, kd, a), 2)
        print(f"n={n}: kern_fwd {ms_k:.2f} ms ({4*n*n/ms_k/1e6:.0f} GB/s lower) | potrf {ms_p:.1f} ms ({n**3/3/ms_p/1e9:.2f} TF) info={info.item()} | trsv {ms_t:.2f} ms ({4*n*n/ms_t/1e6:.0f} GB/s) | potri {ms_i:.1f} ms ({2*n**3/3/ms_i/1e9:.2f} TF) | gpr_grad {ms_g:.2f} ms ({4*n*n/ms_g/1e6:.0f} GB/s)", flush=True)
print("ALL OK" if ok else "SOME BAD")
