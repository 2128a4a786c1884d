"""Feasibility probe for DESIGN.md section 8 ("beyond the FP64 pipe"): FP64-equivalent GEMM by integer slicing
(Ozaki scheme) on B200's INT8 tensor path.  NOT part of the library and not on any product path: the slice products go
through torch._int_mm (cuBLASLt), the split and the recombination are plain torch ops.  It answers three questions
with measurements instead of estimates:

  1. what int8 GEMM rate does this B200 deliver at the shapes of the trailing updates (m = n, k = panel width)?
  2. how many 7-bit slices does C = A B^T need to match a DGEMM, on data conditioned like a covariance matrix?
  3. what is left of the speed-up once the split and the (unfused) recombination are paid?

    python tools/ozaki_probe.py [m] [k]          (defaults 8192 2048)
"""
import json
import sys

import numpy as np
import torch

BETA = 7   # bits per slice; every slice lies in [-64, 64], so a k-long dot product stays below 2^31 for k < 2^19


def split(A, slices):
    """Row-scaled error-free split: A = 2^e[:, None] * sum_s Q_s 2^(-BETA (s + 1)), Q_s int8 in [-64, 64]."""
    amax = A.abs().amax(dim=1, keepdim=True).clamp_min(torch.finfo(A.dtype).tiny)
    e = torch.floor(torch.log2(amax)) + 2.0                  # |A| 2^-e <= 1/2
    r = A * torch.exp2(-e)
    qs = []
    for _ in range(slices):
        r = r * float(2 ** BETA)
        q = torch.round(r)
        r = r - q                                            # exact: |r| <= 1/2 again
        qs.append(q.to(torch.int8))
    return qs, e


def ozaki_matmul_nt(A, B, slices):
    """A (m, k), B (n, k) fp64 -> A B^T with all slice pairs s + t < slices."""
    qa, ea = split(A, slices)
    qb, eb = split(B, slices)
    qbt = [q.t().contiguous() for q in qb]                   # torch._int_mm wants (m, k) x (k, n)
    C = torch.zeros(A.shape[0], B.shape[0], dtype=torch.float64, device=A.device)
    n_prod = 0
    for u in range(slices - 1, -1, -1):                      # smallest weight class first
        acc = None
        for s in range(u + 1):
            p = torch._int_mm(qa[s], qbt[u - s])             # exact int32
            acc = p if acc is None else acc.add_(p)          # still exact: (u + 1) 2^12 k < 2^31
            n_prod += 1
        C.add_(acc.to(torch.float64), alpha=float(2.0 ** (-BETA * (u + 2))))
    C.mul_(torch.exp2(ea)).mul_(torch.exp2(eb).t())
    return C, n_prod


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))[reps // 2]


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(7)
    # operands shaped like a Cholesky trailing update: columns of the factor of an RBF covariance (entries spanning many
    # orders of magnitude along a row), not uniform noise
    X = torch.rand(m, 8, generator=g, dtype=torch.float64)
    Kmat = torch.exp(-0.5 * torch.cdist(X[:k], X[:k]).pow(2)) + 0.01 * torch.eye(k, dtype=torch.float64)
    Lk = torch.linalg.cholesky(Kmat)
    A = (torch.exp(-0.5 * torch.cdist(X, X[:k]).pow(2)) @ torch.linalg.inv(Lk).t()).to(dev)     # rows of L21-like panels
    B = A.clone()
    flop = 2.0 * m * m * k
    out = {"m": m, "n": m, "k": k, "beta_bits": BETA}

    # 1. raw rates
    a8 = torch.randint(-64, 65, (m, k), dtype=torch.int8, device=dev)
    b8 = torch.randint(-64, 65, (k, m), dtype=torch.int8, device=dev)
    ms_i8 = timed(lambda: torch._int_mm(a8, b8))
    ms_f64 = timed(lambda: torch.matmul(A, B.t()))
    out["int8_gemm_ms"] = ms_i8
    out["int8_gemm_tops"] = flop / ms_i8 * 1e-9
    out["cublas_dgemm_ms"] = ms_f64
    out["cublas_dgemm_tflops"] = flop / ms_f64 * 1e-9

    # 2. accuracy against extended precision on a sub-block
    rows = np.arange(0, m, m // 48)[:48]
    cols = np.arange(3, m, m // 48)[:48]
    Ah, Bh = A[rows].cpu().numpy().astype(np.longdouble), B[cols].cpu().numpy().astype(np.longdouble)
    ref = Ah @ Bh.T
    scale = (np.abs(Ah) @ np.abs(Bh).T).astype(np.float64)          # component-wise error scale sum_k |a_ik||b_jk|
    nrm = float(np.abs(ref).max())

    def errs(C):
        d = np.abs(C[rows][:, cols].cpu().numpy().astype(np.longdouble) - ref).astype(np.float64)
        return {"max_abs_over_max_ref": float(d.max() / nrm), "max_componentwise": float((d / scale).max())}

    out["dgemm_error"] = errs(torch.matmul(A, B.t()))
    out["ozaki"] = {}
    for s in (5, 6, 7, 8, 9):
        C, n_prod = ozaki_matmul_nt(A, B, s)
        ms_all = timed(lambda: ozaki_matmul_nt(A, B, s), reps=3)
        ms_split = timed(lambda: (split(A, s), split(B, s)), reps=3)
        out["ozaki"][str(s)] = {"products": n_prod, "error": errs(C), "ms_total_unfused": ms_all, "ms_split": ms_split,
                                "ms_products_only": n_prod * ms_i8,
                                "fp64_equiv_tflops_products_only": flop / (n_prod * ms_i8) * 1e-9,
                                "fp64_equiv_tflops_unfused": flop / ms_all * 1e-9}
        del C
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
