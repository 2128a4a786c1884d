"""Mean functions (reference: gptorch/mean_functions.py); any torch.nn.Module mapping [n, dx] -> [n, dy] works."""
import torch

from . import settings
from .util import torch_dtype


class Constant(torch.nn.Module):
    """m(x) = val, one value per output dimension."""

    def __init__(self, dy, val=None):
        super().__init__()
        if val is None:
            val = torch.zeros(dy, dtype=torch_dtype, device=settings.default_device())
        else:
            if val.shape[0] != dy:
                raise ValueError("Provided val doesn't match output dimension")
            val = val.clone()
        self._dy = dy
        self.val = torch.nn.Parameter(val)

    def forward(self, x):
        return torch.zeros(x.shape[0], self._dy, dtype=torch_dtype, device=self.val.device) + self.val

    def _is_cuda(self):
        return self.val.is_cuda


class Zero(Constant):
    """m(x) = 0 (the default)."""

    def __init__(self, dy):
        super().__init__(dy)
        self.val.requires_grad_(False)
