"""Numerical primitives with the reference's names (gptorch/functions.py), running on libgpb200.so.

cholesky / jit_op keep the reference's failure-driven jitter schedule: first try un-jittered, then absolute
jitter 1e-10, 1e-9, ..., 1e-1 on the diagonal, finally RuntimeError("Max tries exceeded.")
(gptorch/functions.py:20-43).  Unlike the reference, the jitter matrix is built on the operand's device.
"""
import torch

from . import _autograd as ag
from . import _native as nv


def jit_op(op, x, max_tries=10, verbose=False):
    """Try ``op(x)``; on failure retry with growing absolute jitter on the diagonal."""
    try:
        return op(x)
    except Exception:
        if verbose:
            print("Op {} failed (initial try)".format(getattr(op, "__name__", op)))
    for i in range(max_tries):
        try:
            jitter = 10.0 ** (-max_tries + i) * torch.eye(*x.shape, dtype=x.dtype, device=x.device)
            return op(x + jitter)
        except RuntimeError:
            if verbose:
                print("Op {} failed (try {} / {})".format(getattr(op, "__name__", op), i + 1, max_tries))
    raise RuntimeError("Max tries exceeded.")


def _potrf(x):
    """Plain lower Cholesky on the native blocked kernel; raises torch.linalg.LinAlgError if not PD."""
    if x.dim() != 2 or x.shape[0] != x.shape[1]:
        raise ValueError("cholesky expects a square matrix")
    L, dinv = ag.CholeskyFn.apply(x)
    L._gpb_dinv = dinv  # lets trtrs() reuse the diagonal-block inverses
    return L


def cholesky(x):
    """Lower Cholesky factor with jitter retry (gptorch/functions.py:46-47)."""
    return jit_op(_potrf, x)


jitchol = cholesky  # name used by BASELINE.json's north_star


def cholesky_inverse(x, upper=False):
    """(L L^T)^-1 from a Cholesky factor (gptorch/functions.py:50-54); differentiable in the factor."""
    L = x.t() if upper else x
    if L.dim() != 2 or L.shape[0] != L.shape[1]:
        raise ValueError("cholesky_inverse expects a square matrix")
    return ag.CholeskyInverseFn.apply(L)


def inverse(x):
    """Jittered inverse of an SPD matrix (gptorch/functions.py:57-58), via Cholesky."""
    return jit_op(lambda a: cholesky_inverse(_potrf(a)), x)


def lt_log_determinant(L):
    """sum(log(diag(L))) (gptorch/functions.py:61-68)."""
    return ag.LogDetFn.apply(L)


def trtrs(b, a, lower=True):
    """Solve a x = b with triangular a (gptorch/functions.py:71-76)."""
    if a.dim() != 2 or b.dim() != 2 or a.shape[0] != a.shape[1] or a.shape[0] != b.shape[0]:
        raise ValueError("trtrs: incompatible shapes {} and {}".format(tuple(b.shape), tuple(a.shape)))
    if lower:
        L, trans, dinv = a, False, ag._dinv_of(a)
    else:
        L, trans = a.t(), True       # U x = b  <=>  (U^T)^T x = b with U^T lower
        L = L.contiguous()
        dinv = ag._dinv_of(L)
    k = b.shape[1]
    if k <= ag.TRSV_MAX_RHS:
        return ag.TrsvFn.apply(b, L, dinv, trans)
    if not trans:
        return ag.TrsmRightFn.apply(b.t(), L, dinv).t()      # (b^T L^-T)^T = L^-1 b
    return ag.TrsmLeftTFn.apply(b, L, dinv)                  # L^-T b, differentiable in b and a


def mm_nt(a, b):
    """a @ b.t() on the native FP64 engine."""
    return ag.GemmFn.apply(nv.GEMM_NT, a, b)


def mm_tn(a, b):
    """a.t() @ b on the native FP64 engine."""
    return ag.GemmFn.apply(nv.GEMM_TN, a, b)


def mm(a, b, b_lower=False):
    """a @ b on the native FP64 engine; b_lower declares b square lower-triangular (its zero half is skipped)."""
    return ag.GemmFn.apply(nv.GEMM_NN, a, b, b_lower)
