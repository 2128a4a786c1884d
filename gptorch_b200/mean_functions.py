"""Mean functions m(x): [n, dx] -> [n, dy] (reference API: gptorch/mean_functions.py:14-50).

Any torch.nn.Module with that signature can be used as a model's mean function; the two below are the ones the
reference ships.  They never look at the values of x, so the mean is produced by broadcasting the parameter vector
over the rows (one device copy) rather than by allocating an [n, dy] matrix of zeros and adding the vector to it.
"""
import torch

from . import settings
from .util import torch_dtype


def _offset_vector(dy, val):
    """Initial value of the offset parameter: zeros on the default device, or a private copy of `val`."""
    if val is None:
        return torch.zeros(dy, dtype=torch_dtype, device=settings.default_device())
    if val.shape[0] != dy:
        raise ValueError("Provided val doesn't match output dimension")
    return val.detach().clone()


class Constant(torch.nn.Module):
    """The same learnable offset for every input: m(x)[i, :] = val, one entry of `val` per output dimension."""

    def __init__(self, dy, val=None):
        super().__init__()
        self.output_dim = int(dy)
        self.val = torch.nn.Parameter(_offset_vector(self.output_dim, val))

    @property
    def _dy(self):
        return self.output_dim

    def forward(self, x):
        # materialised (not a view of the parameter) so that callers may modify the result in place
        return self.val.unsqueeze(0).expand(x.shape[0], self.output_dim).clone()

    def _is_cuda(self):
        """Whether the offset lives on a CUDA device (kept for callers of the reference's helper)."""
        return self.val.device.type == "cuda"


class Zero(Constant):
    """The default mean of every GP model here: identically zero and excluded from training."""

    def __init__(self, dy):
        super().__init__(dy, val=None)
        self.val.requires_grad = False
