"""Likelihoods (reference: gptorch/likelihoods.py).  Only the Gaussian one exists in the reference."""
import abc
from math import pi

import torch
from torch import distributions

from . import settings
from .model import Model, Param
from .settings import DefaultPositiveTransform
from .util import torch_dtype


class Likelihood(Model):
    """p(y | f), factorising over data."""

    def __init__(self):
        super().__init__()

    def predict_mean_variance(self, mean_f, var_f):
        raise NotImplementedError

    def forward(self):
        return None

    @abc.abstractmethod
    def propagate_log(self, qf, targets):
        """E_q(f)[log p(y | f)]."""
        raise NotImplementedError("Implement quadrature fallback")


class Gaussian(Likelihood):
    """Spherical Gaussian noise with variance `variance` (stored as its log)."""

    def __init__(self, variance=1.0):
        super().__init__()
        v = torch.tensor([float(variance)], dtype=torch_dtype, device=settings.default_device())
        self.variance = Param(v, transform=DefaultPositiveTransform())

    def logp(self, F, Y):
        return distributions.Normal(F, torch.sqrt(self.variance.transform())).log_prob(Y)

    def predict_mean_variance(self, mean_f, var_f):
        return mean_f, var_f + self.variance.transform().expand_as(var_f)

    def predict_mean_covariance(self, mean_f, cov_f):
        n = cov_f.shape[0]
        eye = torch.eye(n, dtype=cov_f.dtype, device=cov_f.device)
        return mean_f, cov_f + self.variance.transform() * eye

    def propagate_log(self, qf, targets):
        """-1/2 [ n (log 2 pi + log s2) + (sum (y - mu)^2 + sum var) / s2 ]  (gptorch/likelihoods.py:125-144)."""
        if not isinstance(qf, (distributions.Normal, distributions.MultivariateNormal)):
            raise TypeError("Expect Gaussian q(f)")
        mu, s = qf.loc, qf.variance
        n = targets.nelement()
        if mu.nelement() != n:
            raise ValueError("Targets (%i) and q(f) (%i) have mismatch in size" % (n, mu.nelement()))
        return self.expected_log_density(mu, s, targets)

    def expected_log_density(self, mu, var, targets):
        s2 = self.variance.transform()
        n = targets.nelement()
        log2pi = torch.log(torch.tensor([2.0 * pi], dtype=torch_dtype, device=s2.device))
        return -0.5 * (n * (log2pi + torch.log(s2)) + (torch.sum((targets - mu) ** 2) + var.sum()) / s2)
