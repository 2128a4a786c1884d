"""Exact GP regression (reference: gptorch/models/gpr.py)."""
import torch

from .. import _autograd as ag
from .. import _native as nv
from .. import kernels
from ..functions import cholesky, trtrs, mm, mm_nt
from ..likelihoods import Gaussian
from .base import GPModel


def _native_kind(kernel):
    """Index of the fused native kernel family, or None when the kernel has to be composed in torch."""
    if isinstance(kernel, kernels.Stationary) and kernel._kind is not None and type(kernel).K is kernels.Stationary.K:
        return kernel._kind
    return None


class GPR(GPModel):
    """Gaussian-process regression with a Gaussian likelihood."""

    def __init__(self, x, y, kernel, mean_function=None, likelihood=None, name="gpr"):
        super().__init__(x, y, kernel, likelihood, mean_function, name)

    def log_likelihood(self, x=None, y=None):
        """log p(y | x, theta), shape [1] (Rasmussen & Williams alg. 2.1; gptorch/models/gpr.py:47-67).

        With a stationary kernel and the Gaussian likelihood the whole evaluation is one fused native node
        (covariance build -> Cholesky -> solve -> log-det, analytic backward); Sum / Product trees of native leaves
        (and a bare Linear) use the composite form of the same node.  Kernels with user-defined leaves go through the
        reference's step-by-step form on the native primitives.
        """
        x = x if x is not None else self.X
        y = y if y is not None else self.Y
        if x.shape[0] != y.shape[0]:
            raise ValueError("X and Y must have same # data.")
        n, dy = y.shape
        resid = y - self.mean_function(x)
        kind = _native_kind(self.kernel)
        if kind is not None and isinstance(self.likelihood, Gaussian):
            return ag.GPRLogLikFn.apply(kind, x, resid, self.kernel.length_scales.transform(),
                                        self.kernel.variance.transform(), self.likelihood.variance.transform())
        terms = kernels.sum_of_products(self.kernel) if isinstance(self.likelihood, Gaussian) else None
        if terms is not None:
            spec, params = kernels.composite_spec(terms)
            return ag.GPRCompositeLogLikFn.apply(spec, x.to(torch.float64), resid, self.likelihood.variance.transform(),
                                                 *params)
        L = cholesky(self._compute_kyy(x=x))
        alpha = trtrs(resid, L)
        red_logdet = ag.LogDetFn.apply(L)
        const = -0.5 * dy * n * torch.log(torch.tensor(2.0 * torch.pi, dtype=alpha.dtype, device=alpha.device))
        return (-0.5 * alpha.pow(2).sum() - dy * red_logdet + const).reshape(1)

    def _compute_kyy(self, x=None):
        """K(x, x) + noise * I (gptorch/models/gpr.py:69-86)."""
        x = x if x is not None else self.X
        kind = _native_kind(self.kernel)
        noise = self.likelihood.variance.transform()
        if kind is not None:
            return ag.KernelFn.apply(kind, x.to(torch.float64), None, self.kernel.length_scales.transform(),
                                     self.kernel.variance.transform(), noise)
        terms = kernels.sum_of_products(self.kernel)
        if terms is not None:
            spec, params = kernels.composite_spec(terms)
            return ag.CompositeKernelFn.apply(spec, x.to(torch.float64), None, noise, *params)
        n = x.shape[0]
        return self.kernel.K(x) + noise * torch.eye(n, dtype=x.dtype, device=x.device)

    def _factor(self, x):
        """L = chol(Ky) and V = L^-1 (Y - m(x)).

        The reference re-factorises on every _predict call (gptorch/models/gpr.py:104).  Under torch.no_grad() the
        factor is cached and reused while parameters and data are unchanged (GPModel._memo); with
        autograd enabled it is recomputed so that predictions stay differentiable exactly as in the reference.
        """
        def compute():
            L = cholesky(self._compute_kyy(x=x))
            return L, trtrs(self.Y - self.mean_function(x), L)

        return self._memo("factor", x, compute)

    def _predict(self, x_new, diag=True, x=None):
        """p(f* | y): mean [n*, dy] and variance [n*, dy] (diag) or covariance [n*, n*]
        (gptorch/models/gpr.py:88-117)."""
        x = x if x is not None else self.X
        k_sy = self.kernel.K(x_new, x)                       # [n*, n]
        L, V = self._factor(x)
        At = trtrs(k_sy.t(), L).t()                          # (L^-1 k_ys)^T, solved on the row panel k_sy L^-T
        mean_f = mm(At, V) + self.mean_function(x_new)
        if diag:
            var_f = (self.kernel.Kdiag(x_new) - ag.RowSumSqFn.apply(At))[:, None].expand_as(mean_f)   # fused Kdiag - colsum(A^2)
        else:
            var_f = self.kernel.K(x_new) - mm_nt(At, At)
        return mean_f, var_f
