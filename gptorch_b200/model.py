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
        """Inverse of _get_param_array.  Values are copied into the existing storage (a captured CUDA graph keeps
        reading the parameters at the same addresses)."""
        start = 0
        for p in self._trainable():
            stop = start + p.numel()
            new = torch.as_tensor(np.reshape(param_array[start:stop], p.shape), dtype=torch_dtype)
            if p.data.shape == new.shape and p.data.dtype == new.dtype:
                p.data.copy_(new)
            else:
                p.data = new.to(p.device)
            start = stop

    def _graphable(self):
        """True when loss() is a static sequence of device work (same shapes, no host-side randomness or reads):
        subclasses opt in."""
        return False

    def _finish_eval(self, value, grad):
        print("loss: %s" % value)
        finite = np.isfinite(grad)
        if np.all(finite):
            return float(value), grad.astype(np.float64)
        print("Warning: inf or nan in gradient: replacing with zeros")
        return value, np.where(finite, grad, 0.0).astype(np.float64)

    def _loss_and_grad(self, param_array):
        """f(x), g(x) for scipy.optimize.minimize(jac=True); non-finite gradient entries become 0."""
        from . import settings
        if settings.cuda_graphs and self.__dict__.get("_graph_state") != "failed" and self._graphable():
            out = GraphedEvaluation.run(self, param_array)
            if out is not None:
                return self._finish_eval(*out)
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
        return self._finish_eval(float(packed[0]), packed[1:])

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


class GraphedEvaluation:
    """One loss() + backward() of a model captured into a CUDA graph and replayed per optimiser step (SURVEY 8f
    row 2: the optimiser bridge without per-evaluation host work).

    Per evaluation the host does: one small host-to-device copy (the flat parameter vector), one graph launch, one
    device-to-host copy of [loss, factorisation status, flat gradient].  Everything else -- scattering the vector into
    the Params, transforms, the native kernels, autograd's backward, packing the gradients -- is inside the graph.
    The factorisation status cannot be read by the host during capture, so it is accumulated on the device
    (_native.deferred_info); a non-zero status after a replay means the un-jittered Cholesky failed and the caller
    re-runs that evaluation eagerly, which applies the reference's jitter schedule (gptorch/functions.py:28-43).
    The capture is keyed on the storage addresses of the parameters and the data; if any of them is replaced the
    graph is rebuilt.  Any failure to capture marks the model as not graphable and the eager path is used.
    """

    def __init__(self, model):
        from . import _native as nv
        self.model = model
        named = [(name, p) for name, p in model.named_parameters() if p.requires_grad]
        self.params = [p for _, p in named]
        # Shadow leaves: fresh Params on the SAME storage.  The real Params may still own AccumulateGrad nodes created
        # by an earlier eager backward on the default stream (kept alive by any loss tensor the user still holds);
        # routing gradients towards those during capture synchronises with the default stream and invalidates the
        # capture.  The shadows have no autograd history and exist only inside the graph.
        self.shadows = {}
        for name, p in named:
            q = torch.Tensor._make_subclass(type(p), p.data, True)
            q.__dict__.update({k: v for k, v in p.__dict__.items()})
            self.shadows[name] = q
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.theta = torch.zeros(n, dtype=torch.float64, device=dev)
        self.out = torch.zeros(n + 2, dtype=torch.float64, device=dev)
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)
        self.key = self._key()
        self.theta.copy_(torch.cat([p.detach().reshape(-1).to(torch.float64) for p in self.params]))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up off the default stream: lazy initialisation, allocator pools
            for _ in range(2):
                self._body(nv)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body(nv)
        torch.cuda.synchronize()

    def _key(self):
        m = self.model
        data = [(t.data_ptr(), tuple(t.shape)) for t in (getattr(m, "X", None), getattr(m, "Y", None)) if t is not None]
        return ([(p.data_ptr(), tuple(p.shape)) for p in m._trainable()], data)

    def _body(self, nv):
        offset = 0
        for p in self.params:
            p.data.copy_(self.theta[offset: offset + p.numel()].view(p.shape))
            offset += p.numel()
        self.info.zero_()
        from torch.nn.utils.stateless import _reparametrize_module
        with nv.deferred_info(self.info), _reparametrize_module(self.model, self.shadows):
            loss = self.model.loss()
            grads = torch.autograd.grad(loss.sum(), list(self.shadows.values()), allow_unused=True)
        for p, g in zip(self.params, grads):
            p.grad = g
        pieces = [(g if g is not None else torch.zeros_like(p)).reshape(-1).to(torch.float64) for p, g in zip(self.params, grads)]
        self.out.copy_(torch.cat([loss.detach().reshape(-1)[:1].to(torch.float64), self.info.to(torch.float64)] + pieces))

    @staticmethod
    def run(model, param_array):
        """(loss, grad) through the model's graph, building it on first use; None = use the eager path."""
        state = model.__dict__.get("_graph_eval")
        try:
            if state is None or state.key != state._key():
                model._set_parameters(param_array)      # shapes / dtypes as the optimiser sees them
                state = GraphedEvaluation(model)
                model.__dict__["_graph_eval"] = state
            state.theta.copy_(torch.as_tensor(np.ascontiguousarray(param_array, dtype=np.float64)))
            state.graph.replay()
            packed = state.out.cpu().numpy()
        except Exception as exc:      # capture is an optimisation: never let it take the evaluation down
            model.__dict__["_graph_state"] = "failed"
            model.__dict__.pop("_graph_eval", None)
            warn("CUDA-graph capture of %s.loss() failed (%s: %s); using eager evaluation"
                 % (type(model).__name__, type(exc).__name__, exc))
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
            return None
        if packed[1] != 0.0 or not np.isfinite(packed[0]):
            return None                # not positive-definite without jitter: the eager path retries with jitter
        return float(packed[0]), packed[2:]
