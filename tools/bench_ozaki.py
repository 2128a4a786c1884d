"""Dev tool for the EXPERIMENTAL int8-sliced GEMM (gpb_gemm_ozaki_nt): accuracy against DGEMM, rate of the SYRK-shaped
trailing update against the DMMA engine, and gpb_potrf_lower / gpb_potri_lower with the path switched on (gpb_ozaki_config).

    python tools/bench_ozaki.py [1] [2] [3] [4]      (sections; default all)
"""
import json
import sys
import torch
sys.path.insert(0, ".")
from gptorch_b200 import _native as nv

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
out = {}


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]



def section1():
    # 1. accuracy on covariance-like operands
    m, n, k = 4096, 2048, 1024
    X = torch.rand(m, 8, generator=g, dtype=torch.float64)
    A = torch.exp(-0.5 * torch.cdist(X, X[:k]).pow(2)).to(dev) * torch.exp(3 * torch.randn(m, 1, generator=g, dtype=torch.float64)).to(dev)
    B = torch.exp(-0.5 * torch.cdist(X[:n] + 0.1, X[:k]).pow(2)).to(dev)
    ref = A @ B.t()
    scale = (A.abs() @ B.abs().t())
    acc = {}
    for s in (4, 6, 7, 8, 9):
        C = nv.gemm_ozaki_nt(A, B, slices=s)
        d = (C - ref).abs()
        acc[s] = {"max_abs_over_max_ref": float(d.max() / ref.abs().max()), "max_componentwise": float((d / scale).max())}
    out["accuracy_vs_dgemm_4096x2048x1024"] = acc
    C0 = torch.randn(m, n, generator=g, dtype=torch.float64).to(dev)
    C1 = nv.gemm_ozaki_nt(A, B, slices=8, alpha=-1.0, beta=1.0, C=C0.clone())
    out["alpha_beta_err"] = float((C1 - (C0 - ref)).abs().max() / ref.abs().max())
    # lower: only blocks on/below the diagonal are written
    Cl = torch.full((m, m), 7.0, dtype=torch.float64, device=dev)
    nv.gemm_ozaki_nt(A, None, slices=8, C=Cl, lower_only=True)
    full = A @ A.t()
    blk = torch.arange(m, device=dev) // 128
    mask = blk[None, :] <= blk[:, None]
    out["lower_err"] = float(((Cl - full).abs() * mask).max() / full.abs().max())
    out["lower_untouched_ok"] = bool((Cl[~mask] == 7.0).all())
    print(json.dumps(out), flush=True)


def section2():
    # 2. rate of the trailing-update shape
    for m, k in ((16384, 2048), (30720, 2048), (30720, 1024)):
        P = torch.randn(m, k, generator=g, dtype=torch.float64).to(dev)
        C, ld = nv._aligned_empty(m, m, dev)
        C.zero_()
        flop = float(m) * m * k          # lower half of 2 m^2 k
        t_d = timed(lambda: nv.gemm(nv.GEMM_NT, P, P, alpha=-1.0, beta=1.0, C=C, lower_only=True))
        res = {"dmma_ms": t_d, "dmma_tflops": flop / t_d * 1e-9}
        for s in (7, 8):
            t_o = timed(lambda: nv.gemm_ozaki_nt(P, None, slices=s, alpha=-1.0, beta=1.0, C=C, lower_only=True))
            res["ozaki%d_ms" % s] = t_o
            res["ozaki%d_equiv_tflops" % s] = flop / t_o * 1e-9
        out["syrk_lower_m%d_k%d" % (m, k)] = res
        print(json.dumps({"syrk_lower_m%d_k%d" % (m, k): res}), flush=True)
        del P, C


def section3():
    # 3. the blocked Cholesky with the path on
    for nn in (16384, 32768):
        X = torch.rand(nn, 8, generator=g, dtype=torch.float64).to(dev)
        ell = torch.ones(8, dtype=torch.float64, device=dev)
        s2 = torch.ones(1, dtype=torch.float64, device=dev)
        noise = torch.full((1,), 0.01, dtype=torch.float64, device=dev)
        buf, ld = nv._aligned_empty(nn, nn, dev)
        res = {}
        Ls = {}
        for s in (0, 8, 7):
            nv.ozaki_config(s)

            def run():
                nv.kern_fwd(0, X, None, ell, s2, noise=noise, lower=True, out=buf, ldk=ld)
                return nv.potrf_(buf, ld)
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nv.kern_fwd(0, X, None, ell, s2, noise=noise, lower=True, out=buf, ldk=ld)
            e0.record()
            dinv, info = nv.potrf_(buf, ld)
            e1.record()
            torch.cuda.synchronize()
            logdet = float(torch.log(torch.diagonal(buf[:, :nn])).sum())
            res["slices%d" % s] = {"potrf_ms": e0.elapsed_time(e1), "info": int(info.item()), "half_logdet": logdet}
            Ls[s] = torch.diagonal(buf[:, :nn]).clone()
        nv.ozaki_config(0)
        for s in (8, 7):
            res["slices%d" % s]["logdet_rel_vs_dmma"] = abs(res["slices%d" % s]["half_logdet"] - res["slices0"]["half_logdet"]) / abs(res["slices0"]["half_logdet"])
            res["slices%d" % s]["diagL_max_rel"] = float(((Ls[s] - Ls[0]).abs() / Ls[0].abs()).max())
        out["potrf_n%d" % nn] = res
        print(json.dumps({"potrf_n%d" % nn: res}), flush=True)
        del buf


def section4():
    # 4. blocked inverse (trtri + lauum) with the path on, against the DMMA result
    for nn in (16384, 32768):
        X = torch.rand(nn, 8, generator=g, dtype=torch.float64).to(dev)
        ell = torch.ones(8, dtype=torch.float64, device=dev)
        s2 = torch.ones(1, dtype=torch.float64, device=dev)
        noise = torch.full((1,), 0.01, dtype=torch.float64, device=dev)
        buf, ld = nv._aligned_empty(nn, nn, dev)
        res, ref_low, ref_kd = {}, None, None
        for s in (0, 8, 7):
            ms = None
            for rep in range(2):                                  # first pass warms the workspace allocations
                nv.ozaki_config(0)
                nv.kern_fwd(0, X, None, ell, s2, noise=noise, lower=True, out=buf, ldk=ld)
                dinv, info = nv.potrf_(buf, ld)                   # the same DMMA factor for every variant
                nv.ozaki_config(s)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                kd = nv.potri_(buf, ld, dinv)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
            nv.ozaki_config(0)
            blk = torch.arange(nn, device=dev) // 128
            r = {"potri_ms": ms}
            low = buf[:, :nn]
            if s == 0:
                ref_kd = kd.clone()
                ref_rows = low[nn - 3000:].clone()                # a band of rows (all columns) as the reference sample
                scale = float(ref_kd.abs().max())
            else:
                mask = (blk[None, :] < blk[nn - 3000:, None])
                r["kd_max_abs_over_max"] = float((kd - ref_kd).abs().max() / scale)
                r["lower_band_max_abs_over_max"] = float(((low[nn - 3000:] - ref_rows).abs() * mask).max() / scale)
            res["slices%d" % s] = r
        out["potri_n%d" % nn] = res
        print(json.dumps({"potri_n%d" % nn: res}), flush=True)
        del buf, ref_rows, ref_kd


if __name__ == "__main__":
    todo = [a for a in sys.argv[1:] if a in "1234"] or ["1", "2", "3", "4"]
    for t in todo:
        {"1": section1, "2": section2, "3": section3, "4": section4}[t]()
