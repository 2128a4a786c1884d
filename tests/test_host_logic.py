"""Host-side mirror of the reference API: Params, Model plumbing, shapes/errors, and the loud failure when a
numerical call is attempted without CUDA.  CPU only (no kernels are launched)."""
import numpy as np
import pytest
import torch

from gptorch_b200 import kernels, likelihoods, mean_functions, settings, util
from gptorch_b200._lib import NativeLibraryError
from gptorch_b200.model import Model
from gptorch_b200.models import GPR, VFE
from gptorch_b200.param import Param
from torch.distributions.transforms import ExpTransform


@pytest.fixture(autouse=True)
def cpu_default_device():
    settings.set_default_device("cpu")
    yield
    settings.set_default_device(None)


def test_param_stores_inverse_transform():
    """test/test_param.py:29-54."""
    p = Param(torch.tensor([2.0, 3.0], dtype=torch.float64), transform=ExpTransform())
    assert torch.allclose(p.data, torch.log(torch.tensor([2.0, 3.0], dtype=torch.float64)))
    assert torch.allclose(p.transform(), torch.tensor([2.0, 3.0], dtype=torch.float64))
    q = Param(torch.tensor([1.5], dtype=torch.float64))
    assert q.transform().item() == 1.5 and q.prior is None


def test_kernel_hyperparameters_are_log_space():
    k = kernels.Rbf(3, ARD=True, length_scales=np.array([0.25, 0.5, 0.75]), variance=2.0)
    assert np.allclose(k.length_scales.data.numpy(), np.log([0.25, 0.5, 0.75]))
    assert np.allclose(k.variance.transform().detach().numpy(), [2.0])
    assert kernels.Matern32(2).length_scales.shape == (1,)
    assert kernels.Linear(4).variance.shape == (4,)
    with pytest.raises(ValueError):
        kernels.Rbf(2) + kernels.Rbf(3)
    s = kernels.Rbf(2) + kernels.Matern52(2)
    assert isinstance(s, kernels.Sum) and isinstance(kernels.Rbf(2) * kernels.Exp(2), kernels.Product)
    assert kernels.Matern12 is not kernels.Exp and issubclass(kernels.Matern12, kernels.Exp)
    assert kernels.SquaredExponential is kernels.Rbf


def test_kdiag_and_static_kernels_need_no_gpu():
    x = torch.rand(5, 2, dtype=torch.float64)
    assert torch.equal(kernels.Rbf(2, variance=1.5).Kdiag(x), torch.full((5,), 1.5, dtype=torch.float64))
    assert kernels.Constant(2, variance=0.5).K(x).shape == (5, 5)
    assert torch.equal(kernels.White(2).K(x, x[:3]), torch.zeros(5, 3, dtype=torch.float64))
    assert torch.allclose(kernels.Linear(2).Kdiag(x), (x * x).sum(1))


def test_numerical_calls_fail_loudly_without_cuda():
    x = torch.rand(6, 2, dtype=torch.float64)
    with pytest.raises(NativeLibraryError, match="no CPU path"):
        kernels.Rbf(2).K(x)
    model = GPR(np.random.rand(8, 2), np.random.rand(8, 1), kernels.Rbf(2))
    with pytest.raises(NativeLibraryError):
        model.loss()
    with pytest.raises(NativeLibraryError):
        model.predict_f(np.random.rand(3, 2))


def test_model_param_array_roundtrip_and_repr():
    """test/test_model.py:55-115 on a real GPR (no compute needed)."""
    model = GPR(np.random.rand(8, 2), np.random.rand(8, 1), kernels.Rbf(2, ARD=True), likelihood=likelihoods.Gaussian(0.01))
    arr = model._get_param_array()
    assert arr.shape == (4,) and np.isclose(arr[-1], np.log(0.01))
    model._set_parameters(np.array([0.1, 0.2, 0.3, 0.4]))
    assert np.allclose(model._get_param_array(), [0.1, 0.2, 0.3, 0.4])
    assert np.allclose(model.kernel.length_scales.transform().detach().numpy(), np.exp([0.2, 0.3]))
    text = repr(model)
    assert "variance" in text and "length_scales" in text and "(kernel)" in text
    assert model.__class__.__name__ == "gpr"          # GPModel renames the class (gptorch/models/base.py:87)
    assert model.num_data == 8 and model.input_dimension == 2 and model.output_dimension == 1
    assert len(model.extract_params()) == 4


def test_log_prior_and_loss_dispatch():
    class Toy(Model):
        def __init__(self):
            super().__init__()
            self.a = Param(torch.tensor([2.0], dtype=torch.float64), transform=ExpTransform(),
                           prior=torch.distributions.Gamma(torch.tensor([2.0], dtype=torch.float64), torch.tensor([1.0], dtype=torch.float64)))

        def _loss(self, shift=0.0):
            return -(self.log_prior()) + shift

    m = Toy()
    expected = torch.distributions.Gamma(torch.tensor([2.0], dtype=torch.float64), torch.tensor([1.0], dtype=torch.float64)).log_prob(torch.tensor([2.0], dtype=torch.float64)).sum()
    assert torch.isclose(m.log_prior(), expected)
    assert torch.isclose(m.loss(shift=1.0), -expected + 1.0)
    assert m.compute_loss is not None
    f, g = m._loss_and_grad(m._get_param_array())
    assert np.isfinite(f) and g.shape == (1,)


def test_shape_mismatch_raises_value_error_before_any_compute():
    model = GPR(np.random.rand(8, 2), np.random.rand(8, 1), kernels.Rbf(2))
    with pytest.raises(ValueError, match="X and Y must have same # data."):
        model.loss(x=model.X[:4], y=model.Y)
    vfe = VFE(np.random.rand(30, 2), np.random.rand(30, 1), kernels.Rbf(2), inducing_points=np.random.rand(4, 2))
    assert vfe.num_inducing == 4 and isinstance(vfe.Z, Param) and vfe.Z.requires_grad
    with pytest.raises(ValueError):
        vfe.loss(x=vfe.X[:3])
    vfe2 = VFE(np.random.rand(55, 2), np.random.rand(55, 1), kernels.Rbf(2))   # default: clip(N // 10, 1, 100) k-means centres
    assert vfe2.num_inducing == 5
    with pytest.raises(AssertionError):
        VFE(np.random.rand(9, 2), np.random.rand(9, 1), kernels.Rbf(2), mean_function=mean_functions.Constant(1))


def test_squared_distance_first_and_second_derivatives():
    """test/test_util.py:25-106 on the package's own composite squared_distance."""
    x1 = torch.tensor([[0.0], [1.0], [2.0]], dtype=torch.float64)
    x2 = torch.tensor([[0.0], [2.0], [4.0]], dtype=torch.float64)
    assert np.array_equal(util.squared_distance(x1, x2).numpy(), [[0, 4, 16], [1, 1, 9], [4, 0, 4]])
    a = torch.tensor([[1.0]], dtype=torch.float64, requires_grad=True)
    (g,) = torch.autograd.grad(util.squared_distance(a, torch.tensor([[1.0]], dtype=torch.float64)).sum(), a, create_graph=True)
    (h,) = torch.autograd.grad(g.sum(), a)
    assert g.item() == 0.0 and h.item() == 2.0


def test_likelihood_and_mean_functions():
    """test/test_likelihoods.py:45-114, test/test_mean_functions.py:12-44."""
    lik = likelihoods.Gaussian(variance=0.01)
    logp = lik.logp(torch.zeros(1, dtype=torch.float64), torch.tensor([0.1], dtype=torch.float64))
    assert logp.item() == pytest.approx(0.8836465597893728)
    mu, var = lik.predict_mean_variance(torch.zeros(3, 1, dtype=torch.float64), torch.ones(3, 1, dtype=torch.float64))
    assert torch.allclose(var, torch.full((3, 1), 1.01, dtype=torch.float64))
    _, cov = lik.predict_mean_covariance(torch.zeros(3, 1, dtype=torch.float64), torch.zeros(3, 3, dtype=torch.float64))
    assert torch.allclose(cov, 0.01 * torch.eye(3, dtype=torch.float64))
    q = torch.distributions.Normal(torch.zeros(4, dtype=torch.float64), torch.ones(4, dtype=torch.float64))
    val = lik.propagate_log(q, torch.full((4,), 0.1, dtype=torch.float64))
    expect = -0.5 * (4 * (np.log(2 * np.pi) + np.log(0.01)) + (4 * 0.01 + 4.0) / 0.01)
    assert val.item() == pytest.approx(expect)
    with pytest.raises(TypeError):
        lik.propagate_log(torch.distributions.Gamma(1.0, 1.0), torch.zeros(1))
    c = mean_functions.Constant(2, val=torch.tensor([1.0, -1.0], dtype=torch.float64))
    assert torch.equal(c(torch.zeros(3, 5)), torch.tensor([[1.0, -1.0]] * 3, dtype=torch.float64))
    z = mean_functions.Zero(2)
    assert not z.val.requires_grad and torch.equal(z(torch.zeros(3, 5)), torch.zeros(3, 2, dtype=torch.float64))
    with pytest.raises(ValueError):
        mean_functions.Constant(3, val=torch.zeros(2, dtype=torch.float64))


def test_importing_package_keeps_torch_default_dtype():
    """test/test_base.py:10-22."""
    assert torch.get_default_dtype() == torch.float32
    assert util.torch_dtype == torch.float64 and util.TensorType is torch.DoubleTensor
    assert util.as_tensor(np.zeros((2, 2), dtype=np.float32)).dtype == torch.float64
    with pytest.raises(TypeError):
        util.as_tensor("nope")


def test_host_batch_stream_sequence_matches_the_rng_stream():
    """HostBatchStream (minibatches of host-resident data) hands out, in order, exactly the rows the
    host RNG draws -- one draw per batch -- while gathering the next batch on a worker thread (CPU device: no copies)."""
    import numpy as np
    from gptorch_b200.models.sparse_gpr import HostBatchStream, draw_minibatch_indices
    rng = np.random.RandomState(0)
    X = torch.as_tensor(rng.rand(500, 3))
    Y = torch.as_tensor(rng.rand(500, 2))
    np.random.seed(11)
    expect = [draw_minibatch_indices(500, 64) for _ in range(5)]
    np.random.seed(11)
    stream = HostBatchStream(X, Y, 64, torch.device("cpu"))
    for idx in expect[:4]:                      # the stream is always one draw ahead
        x, y = stream.next()
        assert torch.equal(x, X[idx]) and torch.equal(y, Y[idx])
    assert len({tuple(i) for i in map(tuple, expect)}) == 5
    # large-N draw: O(batch), without replacement
    from gptorch_b200.models import sparse_gpr as sg
    big = draw_minibatch_indices(sg.MINIBATCH_PERMUTE_MAX + 5, 1000)
    assert len(set(big.tolist())) == 1000 and big.max() < sg.MINIBATCH_PERMUTE_MAX + 5


def test_prediction_memo_key_sees_layout_structure_and_data_changes():
    """GPModel._memo (the factor cache of _predict) must recompute when the view layout, a sub-module, the parameter
    values or the data change -- including edits through `.data` that bump no version counter -- and must hit only for
    the very same tensors in the very same state."""
    model = GPR(np.random.rand(40, 2), np.random.rand(40, 1), kernels.Rbf(2), likelihood=likelihoods.Gaussian(0.1))
    calls = []

    def compute():
        calls.append(1)
        return len(calls)

    X = model.X
    with torch.no_grad():
        assert model._memo("k", X, compute) == 1
        assert model._memo("k", X, compute) == 1                       # hit: same tensors, same state
        a, b = X[:10], X[::2][:10]                                     # same data_ptr and shape, different strides
        assert a.data_ptr() == b.data_ptr() and a.shape == b.shape
        assert model._memo("k", a, compute) == 2
        assert model._memo("k", b, compute) == 3
        assert model._memo("k", X, compute) == 4
        model.kernel = kernels.Matern32(2)                             # same parameter values, different module
        assert model._memo("k", X, compute) == 5
        assert model._memo("k", X, compute) == 5
        model.kernel.variance.data.add_(0.5)                           # parameter edit through .data
        assert model._memo("k", X, compute) == 6
        model.Y = model.Y.clone()                                      # new target tensor with equal values
        assert model._memo("k", X, compute) == 7
        model.X.add_(1.0)                                              # in-place edit (bumps the version counter)
        assert model._memo("k", X, compute) == 8
    assert model._memo("k", X, compute) == 9                           # autograd enabled: never cached
    assert model._memo("k", X, compute) == 10
