"""Package-wide settings (reference: gptorch/settings.py:5-7)."""
import torch
from torch.distributions.transforms import ExpTransform

# Positive hyper-parameters are stored as logs and read through exp(), so gradients are w.r.t. the logs.
DefaultPositiveTransform = ExpTransform

_device = None


def default_device():
    """Device new parameters and data are created on: the current CUDA device when there is one.

    The reference builds everything on the CPU and moves with model.cuda(); this package computes only on
    the GPU, so it starts there.  Without CUDA (e.g. the CPU test tier) objects can still be constructed and
    inspected, but any numerical call raises.
    """
    if _device is not None:
        return _device
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def set_default_device(device):
    global _device
    _device = None if device is None else torch.device(device)
