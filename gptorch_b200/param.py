"""Parameters with a constraint transform and an optional prior (reference: gptorch/param.py:13-50)."""
import torch
from torch.distributions.transforms import ComposeTransform


def _as_transform(t):
    return ComposeTransform([]) if t is None else t


class Param(torch.nn.Parameter):
    """torch Parameter that stores the UNCONSTRAINED value.

    ``Param(v, transform=t)`` stores ``t.inv(v)``; ``.transform()`` returns the constrained value ``t(raw)``.
    Optimisers and ``.grad`` therefore see the raw value (log-space for positive quantities).
    """

    def __new__(cls, data=None, requires_grad=True, transform=None, prior=None):
        raw = _as_transform(transform).inv(data)
        return super().__new__(cls, raw, requires_grad=requires_grad)

    def __init__(self, data, requires_grad=True, transform=None, prior=None):
        super().__init__()
        self._transform = _as_transform(transform)
        self.prior = prior

    def transform(self):
        return self._transform(self)

    def __repr__(self):
        return "Parameter containing:" + self.data.__repr__()

    @staticmethod
    def _validate_transform(t):
        return _as_transform(t)
