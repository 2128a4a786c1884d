"""Constrained parameters (reference API: gptorch/param.py:13-50).

A `Param` is a torch Parameter that holds the UNCONSTRAINED number the optimiser moves and remembers the bijection
to the constrained value the model uses: `Param(v, transform=t)` stores t^-1(v) and `.transform()` evaluates t(raw).
Positive quantities (variances, length scales) use exp, so `.grad` is a gradient with respect to the logarithm; a
`prior` (any torch.distributions object) is evaluated on the constrained value by Model.log_prior.
"""
import torch
from torch.distributions.transforms import ComposeTransform

_IDENTITY = ComposeTransform([])


def _bijection(transform):
    """The transform to use: the identity when none was given."""
    return transform if transform is not None else _IDENTITY


class Param(torch.nn.Parameter):
    def __new__(cls, data=None, requires_grad=True, transform=None, prior=None):
        unconstrained = _bijection(transform).inv(data)
        return torch.nn.Parameter.__new__(cls, unconstrained, requires_grad)

    def __init__(self, data, requires_grad=True, transform=None, prior=None):
        torch.nn.Parameter.__init__(self)
        self.prior = prior
        self._transform = _bijection(transform)

    def transform(self):
        """Constrained value t(raw), differentiable with respect to the raw storage."""
        return self._transform(self)

    @staticmethod
    def _validate_transform(t):
        return _bijection(t)

    def __repr__(self):
        return "Parameter containing:" + repr(self.data)
