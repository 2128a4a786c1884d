"""CPU oracle for the gptorch dense-GP hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional form on torch *CPU* float64 tensors, the algorithm the reference
(cics-nd/gptorch v0.3.2) runs for kernels.K/.Kdiag, functions.cholesky/trtrs/lt_log_determinant and the
GPR / VFE / SVGP log_likelihood and _predict paths.  Each function cites the reference file:line it follows.

The arithmetic itself lives in the reference's third-party dependency `torch` (unpinned: requirements.txt:13
"torch>=1"; here torch 2.11.0 CPU -> MKL LAPACK/BLAS), so the oracle calls the same torch CPU ops in the same
order; gradients come from torch autograd exactly as in the reference.  Pinning: `oracle/make_golden.py`
(run in the build container, where /root/reference is importable) asserts this file reproduces the real
reference bit-for-bit on every case it emits and re-checks the reference's own golden vectors
(test/data/kernels/*.npy, test/data/models/sparse_gpr/*.dat, the 8.842242323920674 / 9.534628739243518 loss
pins); the emitted vectors are committed under tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package (gptorch_b200) never does.
"""
import math

import numpy as np
import torch

DTYPE = torch.float64

STATIONARY = ("Rbf", "Exp", "Matern12", "Matern32", "Matern52")


def as_f64(x):
    """gptorch/util.py:15-31 (as_tensor): numpy / tensor -> CPU DoubleTensor."""
    if isinstance(x, torch.Tensor):
        return x.to(DTYPE)
    return torch.as_tensor(np.asarray(x), dtype=DTYPE)


# --------------------------------------------------------------------------------------------------------
# L1 primitives
# --------------------------------------------------------------------------------------------------------
def sqdist(x1, x2=None):
    """gptorch/util.py:73-88: |a|^2 + |b|^2 - 2 a.b, value clamped at 0 with the gradient left un-clamped."""
    if x2 is None:
        x2 = x1
    a2 = x1.pow(2).sum(1, keepdim=True)
    b2 = x2.pow(2).sum(1, keepdim=True)
    r2 = a2 + b2.t() - 2.0 * x1 @ x2.t()
    return r2 - torch.clamp(r2, max=0.0).detach()


def jitter_retry(op, x, max_tries=10):
    """gptorch/functions.py:20-43 (jit_op): plain try, then absolute jitter 1e-10 ... 1e-1 on the diagonal."""
    try:
        return op(x)
    except Exception:
        pass
    for i in range(max_tries):
        try:
            return op(x + 10.0 ** (-max_tries + i) * torch.eye(*x.shape, dtype=x.dtype))
        except RuntimeError:
            pass
    raise RuntimeError("Max tries exceeded.")


def _potrf(x):
    # torch.cholesky (deprecated alias used at gptorch/functions.py:47) == torch.linalg.cholesky (lower)
    return torch.linalg.cholesky(x)


def chol(x):
    """gptorch/functions.py:46-47."""
    return jitter_retry(_potrf, x)


def tri_solve(b, a, lower=True):
    """gptorch/functions.py:71-76 (trtrs): solve a x = b with triangular a."""
    return torch.linalg.solve_triangular(a, b, upper=not lower)


def tri_logdet(L):
    """gptorch/functions.py:61-68."""
    return L.diag().log().sum()


# --------------------------------------------------------------------------------------------------------
# L2 covariance functions.  Hyper-parameters are passed already transformed (positive values).
# --------------------------------------------------------------------------------------------------------
def scaled_r2(X, X2, ell):
    """gptorch/kernels.py:149-159."""
    return sqdist(X / ell) if X2 is None else sqdist(X / ell, X2 / ell)


_EXACT_DIAGONAL = False


class exact_diagonal:
    """Context manager: a CHECKER VARIANT, not the reference.  Inside it the scaled distance of K(X) (X2 None) is
    exactly 0 on the diagonal (value and gradient) instead of sqrt(round-off of |x|^2 + |x|^2 - 2 x.x) ~ 1e-8.
    It isolates the one documented deviation of the CUDA kernels from the reference (Exp/Matern12):
    everything else in the oracle stays the reference's arithmetic.  oracle/exact_witness.py (mpmath) shows which of
    the two diagonals is the mathematically right one."""

    def __enter__(self):
        global _EXACT_DIAGONAL
        self.prev, _EXACT_DIAGONAL = _EXACT_DIAGONAL, True
        return self

    def __exit__(self, *exc):
        global _EXACT_DIAGONAL
        _EXACT_DIAGONAL = self.prev
        return False


def scaled_r(X, X2, ell):
    """gptorch/kernels.py:161-172."""
    r = torch.sqrt(torch.clamp(scaled_r2(X, X2, ell), min=1e-40))
    if _EXACT_DIAGONAL and X2 is None:
        r = r * (1.0 - torch.eye(X.shape[0], dtype=DTYPE))
    return r


def cov(kind, X, X2, ell, variance):
    """Kernel.K.  gptorch/kernels.py: Rbf :220-222, Exp/Matern12 :189-190, Matern32 :198-201,
    Matern52 :205-212, Linear :258-262 (for Linear, `ell` is unused and `variance` is the [D] vector)."""
    if kind in ("Rbf", "SquaredExponential"):
        return variance * torch.exp(-scaled_r2(X, X2, ell) / 2.0)
    if kind in ("Exp", "Matern12"):
        return variance * torch.exp(-scaled_r(X, X2, ell))
    if kind == "Matern32":
        r3 = math.sqrt(3.0) * scaled_r(X, X2, ell)
        return variance * (1.0 + r3) * torch.exp(-r3)
    if kind == "Matern52":
        r = scaled_r(X, X2, ell)
        s5 = math.sqrt(5.0)
        return variance * (1.0 + s5 * r + 5.0 / 3.0 * r * r) * torch.exp(-s5 * r)
    if kind == "Linear":
        return torch.mm(X * variance, (X if X2 is None else X2).t())
    if kind == "Periodic":      # gptorch/kernels.py:228-235
        return variance * torch.cos(scaled_r(X, X2, ell))
    if kind in ("Constant", "Bias"):   # gptorch/kernels.py:96-101
        n1, n2 = X.size(0), (X if X2 is None else X2).size(0)
        return variance.expand(n1, n2)
    if kind == "White":         # gptorch/kernels.py:83-93
        if X2 is None:
            return variance.expand(X.size(0)).diag()
        return torch.zeros(X.size(0), X2.size(0), dtype=DTYPE)
    raise ValueError(kind)


def cov_composite(expr, leaves, X, X2=None):
    """Sum / Product trees (gptorch/kernels.py:286-306).  `expr` is a Python expression over k0, k1, ... using + and *
    (the reference builds the same tree through Kernel.__add__/__mul__, gptorch/kernels.py:36-40); leaves[i] is
    (kind, ell or None, variance) with tensors.  K of a Sum / Product is the element-wise sum / product of the
    children's K, so the expression is evaluated on the leaf matrices."""
    env = {"k%d" % i: cov(kind, X, X2, ell, var) for i, (kind, ell, var) in enumerate(leaves)}
    return eval(expr, {"__builtins__": {}}, env)  # noqa: S307 (test infrastructure; fixed expressions)


def cov_diag(kind, X, variance):
    """Kernel.Kdiag.  gptorch/kernels.py:174-179 (stationary: broadcast variance), :264-265 (Linear)."""
    if kind == "Linear":
        return torch.sum(X * X * variance, 1)
    return variance.expand(X.size(0))


class Hyper:
    """Raw (log-space) hyper-parameters as autograd leaves, like gptorch Params with ExpTransform
    (gptorch/param.py:23-35, gptorch/settings.py:5-7).  Gradients are w.r.t. these raw values."""

    def __init__(self, kind, ell, variance, noise):
        self.kind = kind
        self.raw_ell = torch.log(as_f64(np.atleast_1d(ell))).requires_grad_(True) if ell is not None else None
        self.raw_var = torch.log(as_f64(np.atleast_1d(variance))).requires_grad_(True)
        self.raw_noise = torch.log(as_f64(np.atleast_1d(noise))).requires_grad_(True)

    @property
    def ell(self):
        return None if self.raw_ell is None else torch.exp(self.raw_ell)

    @property
    def variance(self):
        return torch.exp(self.raw_var)

    @property
    def noise(self):
        return torch.exp(self.raw_noise)

    def leaves(self):
        return [p for p in (self.raw_var, self.raw_ell, self.raw_noise) if p is not None]

    def K(self, X, X2=None):
        return cov(self.kind, X, X2, self.ell, self.variance)

    def Kdiag(self, X):
        return cov_diag(self.kind, X, self.variance)


# --------------------------------------------------------------------------------------------------------
# L3 models
# --------------------------------------------------------------------------------------------------------
def gpr_kyy(h, X):
    """gptorch/models/gpr.py:69-86."""
    n = X.shape[0]
    return h.K(X) + h.noise.expand(n, n).diag().diag()


def gpr_loglik(h, X, Y, mean=None):
    """gptorch/models/gpr.py:47-67.  Returns a shape-[1] tensor like the reference."""
    if X.shape[0] != Y.shape[0]:
        raise ValueError("X and Y must have same # data.")
    n, dy = Y.shape
    L = chol(gpr_kyy(h, X))
    resid = Y - (mean(X) if mean is not None else torch.zeros(n, dy, dtype=DTYPE))
    alpha = tri_solve(resid, L)
    const = torch.tensor([-0.5 * dy * n * np.log(2 * np.pi)], dtype=DTYPE)
    return -0.5 * alpha.pow(2).sum() - dy * tri_logdet(L) + const


def gpr_predict(h, X, Y, Xs, diag=True, mean=None):
    """gptorch/models/gpr.py:88-117."""
    zero = lambda x: torch.zeros(x.shape[0], Y.shape[1], dtype=DTYPE)
    m = mean if mean is not None else zero
    k_ys = h.K(X, Xs)
    L = chol(gpr_kyy(h, X))
    A = tri_solve(k_ys, L)
    V = tri_solve(Y - m(X), L)
    mu = A.t() @ V + m(Xs)
    if diag:
        var = (h.Kdiag(Xs) - (A * A).sum(0))[:, None].expand_as(mu)
    else:
        var = h.K(Xs) - A.t() @ A
    return mu, var


def vfe_elbo(h, X, Y, Z):
    """gptorch/models/sparse_gpr.py:108-153 (Titsias bound; zero mean function).  0-dim tensor."""
    if X.shape[0] != Y.shape[0]:
        raise ValueError("X and Y must have same # data.")
    M, N, dy = Z.shape[0], X.shape[0], Y.shape[1]
    kff = h.Kdiag(X)
    Kuf = h.K(Z, X)
    Kuu = h.K(Z)
    L = chol(Kuu)
    A = tri_solve(Kuf, L)
    AAT = A @ A.t() / h.noise.expand_as(Kuu)
    B = AAT + torch.eye(M, dtype=DTYPE)
    LB = chol(B)
    c = tri_solve(A @ Y, LB) / h.noise
    elbo = torch.tensor([-0.5 * dy * N * np.log(2 * np.pi)], dtype=DTYPE)
    elbo = elbo - dy * LB.diag().log().sum()
    elbo = elbo - 0.5 * dy * N * h.noise.log()
    elbo = elbo - 0.5 * (Y.pow(2).sum() + dy * kff.sum()) / h.noise
    elbo = elbo + 0.5 * c.pow(2).sum()
    elbo = elbo + 0.5 * dy * AAT.diag().sum()
    return elbo[0]


def vfe_predict(h, X, Y, Z, Xs, diag=True):
    """gptorch/models/sparse_gpr.py:155-195."""
    M = Z.shape[0]
    Kuf = h.K(Z, X)
    Kuu = h.K(Z)
    Kus = h.K(Z, Xs)
    L = chol(Kuu)
    A = tri_solve(Kuf, L)
    AAT = A @ A.t() / h.noise.expand_as(Kuu)
    B = AAT + torch.eye(M, dtype=DTYPE)
    LB = chol(B)
    c = tri_solve(A @ Y, LB) / h.noise
    t1 = tri_solve(Kus, L)
    t2 = tri_solve(t1, LB)
    mu = t2.t() @ c
    if diag:
        var = (h.Kdiag(Xs) - t1.pow(2).sum(0).squeeze() + t2.pow(2).sum(0).squeeze())[:, None].expand_as(mu)
    else:
        var = h.K(Xs) + t2.t() @ t2 - t1.t() @ t1
    return mu, var


def gaussian_expected_log(noise, mu, var, y):
    """gptorch/likelihoods.py:125-144 (Gaussian.propagate_log) for q(f) = N(mu, var) elementwise."""
    n = y.nelement()
    return -0.5 * (n * (math.log(2.0 * math.pi) + torch.log(noise)) + (torch.sum((y - mu) ** 2) + var.sum()) / noise)


def lower_cholesky_transform(raw):
    """torch.distributions.transforms.LowerCholeskyTransform: strict lower + exp(diagonal)."""
    return raw.tril(-1) + raw.diagonal(dim1=-2, dim2=-1).exp().diag_embed()


def svgp_predict(h, Z, q_mu, q_sqrt, Xs, diag=True, chol_kuu=None):
    """gptorch/models/sparse_gpr.py:337-381 (zero mean function)."""
    Lu = chol(h.K(Z)) if chol_kuu is None else chol_kuu
    kuf = h.K(Z, Xs)
    alpha = tri_solve(kuf, Lu).t()
    beta = tri_solve(q_sqrt, Lu)
    mu = alpha @ tri_solve(q_mu, Lu)
    gamma = alpha @ beta
    if diag:
        var = (h.Kdiag(Xs) - torch.sum(alpha ** 2, dim=1) + torch.sum(gamma ** 2, dim=1))[:, None].expand_as(mu)
    else:
        var = h.K(Xs) - alpha @ alpha.t() + gamma @ gamma.t()
    return mu, var


def svgp_elbo(h, X, Y, Z, q_mu, q_sqrt, num_data):
    """gptorch/models/sparse_gpr.py:263-308 on an explicit (mini)batch X, Y; zero mean function."""
    if X.shape[0] != Y.shape[0]:
        raise ValueError("X and Y must have same # data.")
    Lu = chol(h.K(Z))
    mu, var = svgp_predict(h, Z, q_mu, q_sqrt, X, diag=True, chol_kuu=Lu)
    data = torch.stack([gaussian_expected_log(h.noise, m_i, v_i, y_i) for m_i, v_i, y_i in zip(mu.t(), var.t(), Y.t())]).sum()
    data = data * (num_data / X.shape[0])
    zero_mean = torch.zeros_like(q_mu)
    kl = torch.stack([
        torch.distributions.kl_divergence(
            torch.distributions.MultivariateNormal(qm, scale_tril=q_sqrt),
            torch.distributions.MultivariateNormal(pm, scale_tril=Lu))
        for qm, pm in zip(q_mu.t(), zero_mean.t())]).sum()
    return data - kl


# --------------------------------------------------------------------------------------------------------
# Seeded synthetic inputs (SURVEY.md 8d / BASELINE.md 3) -- shared by the oracle, the tests and bench.py
# --------------------------------------------------------------------------------------------------------
def synth_regression(n, d, seed=1234):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, d, generator=g, dtype=DTYPE)
    w = torch.randn(d, 1, generator=g, dtype=DTYPE)
    Y = torch.sin(X @ w) + 0.1 * torch.randn(n, 1, generator=g, dtype=DTYPE)
    return X, Y, g


def synth_inducing(X, m, g):
    return X[torch.randperm(X.shape[0], generator=g)[:m]].clone()


def gpr_loss_and_grads(kind, X, Y, ell, variance, noise):
    """loss = -log_likelihood (gptorch/models/base.py:418-419, no priors) and d loss / d raw params."""
    h = Hyper(kind, ell, variance, noise)
    loss = -gpr_loglik(h, X, Y)
    loss.sum().backward()
    return loss.detach(), {"variance": h.raw_var.grad.clone(), "length_scales": h.raw_ell.grad.clone(),
                           "noise": h.raw_noise.grad.clone()}


# --------------------------------------------------------------------------------------------------------
# seeded helpers of the large pins (tests/golden/large_cases.npz)
# --------------------------------------------------------------------------------------------------------
def seeded_q(m, dy, seed=4321):
    """q(u) for the SVGP pin: mean ~ 0.3 N(0,1); raw Cholesky factor = strictly-lower 0.02 N(0,1) with a raw
    (log) diagonal of -1 + 0.1 N(0,1).  Used by oracle/make_golden_large.py and by the tests."""
    g = torch.Generator().manual_seed(seed)
    q_mu = 0.3 * torch.randn(m, dy, generator=g, dtype=torch.float64)
    raw = torch.tril(0.02 * torch.randn(m, m, generator=g, dtype=torch.float64), -1)
    raw = raw + torch.diag(-1.0 + 0.1 * torch.randn(m, generator=g, dtype=torch.float64))
    return q_mu, raw


def projections(G, seed=99, k=8):
    g = torch.Generator().manual_seed(seed)
    V = torch.randn(G.shape[1], k, generator=g, dtype=torch.float64).numpy()
    return np.asarray(G) @ V
