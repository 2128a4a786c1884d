"""Covariance functions with the reference's class names and call surface (gptorch/kernels.py).

Stationary kernels (Rbf, Exp/Matern12, Matern32, Matern52, Periodic) and Linear evaluate K through the fused CUDA
kernel (one pass, analytic backward).  Sum / Product trees whose leaves are those kernels, Constant/Bias or White
are rewritten as a sum of products of leaves and evaluated by ONE pass of the composite kernel
(gpb_kern_sop_fwd) instead of one N x N tensor per child; trees with user-defined leaves, or larger
than the native limits, compose tensors like the reference does.
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
    """variance * cos(r)  (gptorch/kernels.py:228-235)"""
    _kind = nv.KIND["Periodic"]


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
        if X.is_cuda:       # gpb_linear_kdiag; host tensors keep the composite (Kdiag of a CPU tensor needs no kernel)
            return ag.LinearKdiagFn.apply(_f64(X), self.variance.transform())
        return torch.sum(X * X * self.variance.transform(), 1)


class Combination(Kernel):
    """A pair of kernels on the same inputs."""

    def __init__(self, kern1, kern2):
        if kern1.input_dim != kern2.input_dim:
            raise ValueError("Kernels need the same input_dim")
        super().__init__(input_dim=kern1.input_dim)
        self.kern1 = kern1
        self.kern2 = kern2


def _leaf_kind(kernel):
    """Native family index of a leaf the composite kernel can evaluate, or None (a subclass that overrides K() is
    user code and is composed through its own K)."""
    cls = type(kernel)
    if isinstance(kernel, Stationary):
        return kernel._kind if (kernel._kind is not None and cls.K is Stationary.K) else None
    if isinstance(kernel, Linear):
        return nv.KERN_LINEAR if cls.K is Linear.K else None
    if isinstance(kernel, Constant):
        return nv.KERN_CONSTANT if cls.K is Constant.K else None
    if isinstance(kernel, White):
        return nv.KERN_WHITE if cls.K is White.K else None
    return None


def sum_of_products(kernel):
    """The kernel as a list of terms, each a list of leaf kernels, such that K = sum_t prod_{l in t} k_l -- or None
    when a leaf is not a native family or the expansion exceeds the composite kernel's limits."""
    if isinstance(kernel, Combination):
        a, b = sum_of_products(kernel.kern1), sum_of_products(kernel.kern2)
        if a is None or b is None:
            return None
        terms = a + b if isinstance(kernel, Sum) else [ta + tb for ta in a for tb in b]
        if not isinstance(kernel, (Sum, Product)) or type(kernel).K not in (Sum.K, Product.K):
            return None
        if len(terms) > nv.SOP_MAX_TERMS or sum(len(t) for t in terms) > nv.SOP_MAX_LEAVES:
            return None
        return terms
    return [[kernel]] if _leaf_kind(kernel) is not None else None


def composite_spec(terms):
    """(spec, params) for the composite autograd nodes: spec mirrors `terms` with (kind, ell index, sigma2 index)
    into the de-duplicated list of transformed parameter tensors."""
    params, index = [], {}

    def slot(param):
        key = id(param)
        if key not in index:
            index[key] = len(params)
            params.append(param.transform())
        return index[key]

    spec = []
    for term in terms:
        row = []
        for leaf in term:
            kind = _leaf_kind(leaf)
            if kind == nv.KERN_LINEAR:
                row.append((kind, slot(leaf.variance), -1))
            elif kind >= nv.KERN_CONSTANT:
                row.append((kind, -1, slot(leaf.variance)))
            else:
                row.append((kind, slot(leaf.length_scales), slot(leaf.variance)))
        spec.append(tuple(row))
    return tuple(spec), params


class _Fusable(Combination):
    """Sum / Product: one fused pass when every leaf is native, the reference's tensor composition otherwise."""

    def K(self, X, X2=None):
        terms = sum_of_products(self)
        if terms is None:
            return self._compose(self.kern1.K(X, X2), self.kern2.K(X, X2))
        spec, params = composite_spec(terms)
        return ag.CompositeKernelFn.apply(spec, _f64(X), None if X2 is None else _f64(X2), None, *params)


class Product(_Fusable):
    @staticmethod
    def _compose(a, b):
        return a * b

    def Kdiag(self, X):
        return self.kern1.Kdiag(X) * self.kern2.Kdiag(X)


class Sum(_Fusable):
    @staticmethod
    def _compose(a, b):
        return a + b

    def Kdiag(self, X):
        return self.kern1.Kdiag(X) + self.kern2.Kdiag(X)
