"""Base Model class (reference: gptorch/model.py): parameter flattening for scipy, priors, loss dispatch."""
from warnings import warn

import numpy as np
import torch
from torch.autograd import gradcheck

from .param import Param
from .util import TensorType, torch_dtype


def _indent_tail(text, spaces):
    lines = text.split("\n")
    if len(lines) == 1:
        return text
    pad = " " * spaces
    return "\n".join([lines[0]] + [pad + ln for ln in lines[1:]])


class Model(torch.nn.Module):
    """torch Module whose parameters are Params (transform + prior)."""

    def forward(self):
        return None

    def __repr__(self):
        out = self.__class__.__name__ + " (\n"
        for name, p in self._parameters.items():
            out += name + "\n" + str(p.transform().data) + "\n"
        for key, module in self._modules.items():
            out += "  (" + key + "): " + _indent_tail(module.__repr__(), 2) + "\n"
        return out + ")" + "\n"

    # ---- scipy.optimize bridge (gptorch/model.py:56-133) ---------------------------------------------
    def _trainable(self):
        return [p for p in self.parameters() if p.requires_grad]

    def _get_param_array(self):
        """All trainable raw parameter values, flattened and concatenated, as a numpy vector."""
        return np.concatenate([p.detach().cpu().numpy().flatten() for p in self._trainable()])

    def _set_parameters(self, param_array):
        """Inverse of _get_param_array."""
        start = 0
        for p in self._trainable():
            stop = start + p.numel()
            new = torch.as_tensor(np.reshape(param_array[start:stop], p.shape), dtype=torch_dtype)
            p.data = new.to(p.device)
            start = stop

    def _loss_and_grad(self, param_array):
        """f(x), g(x) for scipy.optimize.minimize(jac=True); non-finite gradient entries become 0."""
        self._set_parameters(param_array)
        for _, p in self.named_parameters():
            if p.grad is not None:
                p.grad.data.zero_()
        loss = self.loss()
        loss.backward()
        # one device-side concatenation and ONE device-to-host copy for loss and all gradients (the reference
        # copies parameter by parameter, a sync each)
        pieces = [p.grad.reshape(-1).to(torch.float64) for _, p in self.named_parameters() if p.requires_grad]
        packed = torch.cat([loss.detach().reshape(-1)[:1].to(torch.float64)] + pieces).cpu().numpy()
        value, grad = float(packed[0]), packed[1:]
        print("loss: %s" % value)
        finite = np.isfinite(grad)
        if np.all(finite):
            return float(value), grad.astype(np.float64)
        print("Warning: inf or nan in gradient: replacing with zeros")
        return value, np.where(finite, grad, 0.0).astype(np.float64)

    # ---- gradcheck helpers (gptorch/model.py:138-156, 199-217) -----------------------------------------
    def extract_params(self):
        return tuple(self.parameters())

    def expand_params(self, *args):
        for arg, (_, p) in zip(args, self.named_parameters()):
            if isinstance(arg, Param):
                p.data = arg.data
            elif isinstance(arg, np.ndarray):
                raise NotImplementedError("Unresolved issues with expanding numpy arrays")

    def log_prior(self):
        """Sum of prior log-densities over every parameter that has a prior (gptorch/model.py:158-177)."""
        total = 0.0
        for p in self.parameters():
            if getattr(p, "prior", None) is not None:
                value = p.transform() if getattr(p, "transform", None) is not None else p.data
                total += p.prior.log_prob(value).sum()
        return total

    def loss(self, *loss_args, params=None, **loss_kwargs):
        """Loss with the arguments forwarded to the subclass's _loss (gptorch/model.py:179-197)."""
        if params is not None:
            self.expand_params(*params)
        return self._loss(*loss_args, **loss_kwargs)

    # CHANGELOG.md:14-20 of the reference calls this entry point compute_loss
    compute_loss = loss

    def gradcheck(self, eps=1e-6, atol=1e-5, rtol=1e-3, verbose=False):
        if verbose:
            warn("Verbose not yet figured out")
        return gradcheck(self.loss, self.extract_params(), eps=eps, atol=atol, rtol=rtol)

    def _loss(self, *args, **kwargs):
        raise NotImplementedError("Implement loss function")
