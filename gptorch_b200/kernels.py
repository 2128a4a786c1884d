"""Covariance functions with the reference's class names and call surface (gptorch/kernels.py).

Stationary kernels (Rbf, Exp/Matern12, Matern32, Matern52) and Linear evaluate K through the fused CUDA
kernel (one pass, analytic backward); Static kernels, Periodic and the Sum/Product combinators compose
tensors like the reference does.
"""
import numpy as np
import torch

from . import _autograd as ag
from . import _native as nv
from . import settings
from .model import Model
from .param import Param
from .settings import DefaultPositiveTransform
from .util import as_tensor, squared_distance, torch_dtype


def _k_shape(X, X2):
    return (X.size(0),) * 2 if X2 is None else (X.size(0), X2.size(0))


def _f64(t):
    return t if t.dtype == torch_dtype else t.to(torch_dtype)


def _positive(values):
    """Param holding positive values through the default (exp) transform, on the default device."""
    t = torch.as_tensor(np.atleast_1d(np.asarray(values, dtype=np.float64)), dtype=torch_dtype)
    return Param(t.to(settings.default_device()), transform=DefaultPositiveTransform())


class Kernel(Model):
    """Base class: K(X, X2=None) -> [n1, n2], Kdiag(X) -> [n]; `+` and `*` build Sum / Product."""

    def __init__(self, input_dim):
        self.input_dim = int(input_dim)
        super().__init__()

    def __add__(self, other):
        return Sum(self, other)

    def __mul__(self, other):
        return Product(self, other)

    def _validate_ard_shape(self, x, ARD=None):
        if ARD is None:
            ARD = np.asarray(x).squeeze().shape != ()
        x = x * np.ones(self.input_dim)
        if x.shape != (self.input_dim,):
            raise ValueError("shape of possibly-ARD param does not match input_dim")
        return x, ARD


class Static(Kernel):
    """Kernels that do not look at the inputs; one variance parameter (gptorch/kernels.py:67-80)."""

    def __init__(self, input_dim, variance=1.0):
        super().__init__(input_dim)
        self.variance = _positive([variance])

    def Kdiag(self, X):
        return self.variance.transform().expand(X.size(0))


class White(Static):
    def K(self, X, X2=None, presliced=False):
        if X2 is None:
            return self.variance.transform().expand(X.size(0)).diag()
        return torch.zeros(*_k_shape(X, X2), dtype=torch_dtype, device=X.device)


class Constant(Static):
    def K(self, X, X2=None, presliced=False):
        return self.variance.transform().expand(*_k_shape(X, X2))


class Bias(Constant):
    pass


class Stationary(Kernel):
    """k depends on r = |x - x'| / length_scale; ARD gives one length scale per input dimension
    (gptorch/kernels.py:108-179)."""

    _kind = None  # index into the native kernel families

    def __init__(self, input_dim, variance=1.0, length_scales=None, ARD=False):
        super().__init__(input_dim)
        self.variance = _positive([variance])
        self.ARD = ARD
        if ARD:
            if length_scales is None:
                length_scales = np.ones(input_dim)
            elif isinstance(length_scales, np.ndarray):
                assert len(length_scales) == input_dim
            else:
                length_scales = length_scales * np.ones(input_dim)
            self.length_scales = _positive(length_scales)
        else:
            self.length_scales = _positive([1.0 if length_scales is None else length_scales])

    def squared_dist(self, X, X2):
        """Scaled squared distance (composite torch ops; K() itself uses the fused kernel)."""
        ell = self.length_scales.transform()
        return squared_distance(X / ell) if X2 is None else squared_distance(X / ell, X2 / ell)

    def dist(self, X, X2):
        return torch.sqrt(torch.clamp(self.squared_dist(X, X2), min=1e-40))

    def Kdiag(self, X):
        if isinstance(X, np.ndarray):
            X = as_tensor(X)
        return self.variance.transform().expand(X.size(0))

    def K(self, X, X2=None):
        if self._kind is None:
            raise NotImplementedError
        return ag.KernelFn.apply(self._kind, _f64(X), None if X2 is None else _f64(X2),
                                 self.length_scales.transform(), self.variance.transform())


class Exp(Stationary):
    """variance * exp(-r)"""
    _kind = nv.KIND["Exp"]


class Matern12(Exp):
    pass


class Matern32(Stationary):
    """variance * (1 + sqrt(3) r) exp(-sqrt(3) r)"""
    _kind = nv.KIND["Matern32"]


class Matern52(Stationary):
    """variance * (1 + sqrt(5) r + 5/3 r^2) exp(-sqrt(5) r)"""
    _kind = nv.KIND["Matern52"]


class Rbf(Stationary):
    """variance * exp(-r^2 / 2)"""
    _kind = nv.KIND["Rbf"]


SquaredExponential = Rbf


class Periodic(Stationary):
    """variance * cos(r)  (gptorch/kernels.py:228-235); composite torch ops."""

    def K(self, X, X2=None):
        return self.variance.transform() * torch.cos(self.dist(X, X2))


class Linear(Kernel):
    """sum_d v_d x_d x'_d with one variance per input dimension (gptorch/kernels.py:238-265)."""

    def __init__(self, input_dim, variance=1.0, ARD=None):
        super().__init__(input_dim)
        variance, self.ARD = self._validate_ard_shape(variance, ARD)
        self.variance = _positive(variance)

    def K(self, X, X2=None):
        return ag.KernelFn.apply(nv.KERN_LINEAR, _f64(X), None if X2 is None else _f64(X2),
                                 self.variance.transform(), None)

    def Kdiag(self, X):
        return torch.sum(X * X * self.variance.transform(), 1)


class Combination(Kernel):
    """A pair of kernels on the same inputs."""

    def __init__(self, kern1, kern2):
        if kern1.input_dim != kern2.input_dim:
            raise ValueError("Kernels need the same input_dim")
        super().__init__(input_dim=kern1.input_dim)
        self.kern1 = kern1
        self.kern2 = kern2


class Product(Combination):
    def K(self, X, X2=None):
        return self.kern1.K(X, X2) * self.kern2.K(X, X2)

    def Kdiag(self, X):
        return self.kern1.Kdiag(X) * self.kern2.Kdiag(X)


class Sum(Combination):
    def K(self, X, X2=None):
        return self.kern1.K(X, X2) + self.kern2.K(X, X2)

    def Kdiag(self, X):
        return self.kern1.Kdiag(X) + self.kern2.Kdiag(X)
