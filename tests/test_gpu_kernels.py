"""gptorch_b200.kernels on CUDA against the reference's golden files (test/test_kernels.py, 52 .npy fixtures
re-packed in tests/golden/reference_fixtures.npz) and the oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

STATIONARY = ["Exp", "Matern12", "Matern32", "Matern52", "Rbf"]


def _x(fixtures):
    return (torch.as_tensor(fixtures.raw("kern/x1")).cuda(), torch.as_tensor(fixtures.raw("kern/x2")).cuda())


@pytest.mark.parametrize("name", STATIONARY + ["Linear", "Constant", "Bias", "White", "Periodic"])
def test_kernel_goldens(fixtures, name):
    """K(x1), K(x1, x2), K(x2, x1)^T, Kdiag, symmetry (test/test_kernels.py:59-80)."""
    from gptorch_b200 import kernels
    x1, x2 = _x(fixtures)
    kern = getattr(kernels, name)(3)
    kx = kern.K(x1).detach().cpu().numpy()
    kx2 = kern.K(x1, x2).detach().cpu().numpy()
    kx2t = kern.K(x2, x1).detach().cpu().numpy()
    assert np.allclose(fixtures.raw("kern/%s_kx" % name), kx)
    assert np.allclose(fixtures.raw("kern/%s_kx2" % name), kx2)
    assert np.allclose(kx.T, kx)
    assert np.allclose(fixtures.raw("kern/%s_kx2" % name), kx2t.T)
    assert np.allclose(fixtures.raw("kern/%s_kdiag" % name), kern.Kdiag(x1).detach().cpu().numpy())
    # tighter than the reference's allclose for the fused kernels (Exp's diagonal noise floor excepted)
    if name in ("Matern32", "Matern52", "Rbf", "Linear"):
        assert rel_err(kx, fixtures.raw("kern/%s_kx" % name)) < 1e-13
        assert rel_err(kx2, fixtures.raw("kern/%s_kx2" % name)) < 1e-13


@pytest.mark.parametrize("name", STATIONARY + ["Periodic"])
def test_stationary_shift_and_ard(fixtures, name):
    """Shift invariance x + 0.34 and ARD length scales (test/test_kernels.py:83-127)."""
    from gptorch_b200 import kernels
    x1, x2 = _x(fixtures)
    kern = getattr(kernels, name)(3)
    assert np.allclose(fixtures.raw("kern/%s_kx" % name), kern.K(x1 + 0.34).detach().cpu().numpy())
    assert np.allclose(fixtures.raw("kern/%s_kdiag" % name), kern.Kdiag(x1 + 0.34).detach().cpu().numpy())
    ard = getattr(kernels, name)(3, ARD=True, length_scales=fixtures.raw("kern/ard_length_scales").copy())
    assert np.allclose(fixtures.raw("kern/%s_kx_ard" % name), ard.K(x1).detach().cpu().numpy())
    assert np.allclose(fixtures.raw("kern/%s_kx2_ard" % name), ard.K(x1, x2).detach().cpu().numpy())
    assert np.allclose(fixtures.raw("kern/%s_kdiag_ard" % name), ard.Kdiag(x1).detach().cpu().numpy())


@pytest.mark.parametrize("name", STATIONARY + ["Linear", "Constant", "White"])
def test_add_mul_are_bit_exact(fixtures, name):
    """k + k == Sum(k, k) and k * k == Product(k, k) element for element (test/test_kernels.py:39-57): the forward
    kernel must be run-to-run deterministic."""
    from gptorch_b200 import kernels
    x1, _ = _x(fixtures)
    kern = getattr(kernels, name)(3)
    assert torch.equal((kern + kern).K(x1), kernels.Sum(kern, kern).K(x1))
    assert torch.equal((kern * kern).K(x1), kernels.Product(kern, kern).K(x1))
    big = torch.rand(700, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(3)).cuda()
    assert torch.equal(kern.K(big), kern.K(big))


@pytest.mark.parametrize("name", ["Rbf", "Matern32", "Matern52", "Exp"])
@pytest.mark.parametrize("ard", [True, False])
def test_kernel_gradients_match_oracle_autograd(name, ard):
    """dK/d(raw hyper-parameters, X, X2) through the fused backward vs torch autograd on the oracle's composite ops."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels
    g = torch.Generator().manual_seed(11)
    d = 5
    X = torch.rand(150, d, generator=g, dtype=torch.float64)
    X2 = torch.rand(333, d, generator=g, dtype=torch.float64)
    G = torch.randn(150, 333, generator=g, dtype=torch.float64)
    ell = 0.5 + 0.1 * np.arange(d) if ard else 0.8
    kern = getattr(kernels, name)(d, ARD=ard, length_scales=(np.array(ell) if ard else ell), variance=1.3)
    Xc, X2c = X.cuda().requires_grad_(True), X2.cuda().requires_grad_(True)
    (kern.K(Xc, X2c) * G.cuda()).sum().backward()
    raw_ell = torch.log(torch.as_tensor(np.atleast_1d(ell), dtype=torch.float64)).requires_grad_(True)
    raw_var = torch.log(torch.tensor([1.3], dtype=torch.float64)).requires_grad_(True)
    Xo, X2o = X.clone().requires_grad_(True), X2.clone().requires_grad_(True)
    (O.cov(name, Xo, X2o, raw_ell.exp(), raw_var.exp()) * G).sum().backward()
    tol = 1e-9 if name != "Exp" else 1e-6
    assert rel_err(kern.length_scales.grad.cpu().numpy(), raw_ell.grad.numpy()) < tol
    assert rel_err(kern.variance.grad.cpu().numpy(), raw_var.grad.numpy()) < tol
    assert rel_err(Xc.grad.cpu().numpy(), Xo.grad.numpy()) < tol
    assert rel_err(X2c.grad.cpu().numpy(), X2o.grad.numpy()) < tol
    # symmetric form K(X): both argument roles contribute to dX
    kern.zero_grad()
    Gs = torch.randn(150, 150, generator=g, dtype=torch.float64)
    Xc2 = X.cuda().requires_grad_(True)
    (kern.K(Xc2) * Gs.cuda()).sum().backward()
    Xo2 = X.clone().requires_grad_(True)
    (O.cov(name, Xo2, None, raw_ell.exp().detach(), raw_var.exp().detach()) * Gs).sum().backward()
    assert rel_err(Xc2.grad.cpu().numpy(), Xo2.grad.numpy()) < max(tol, 1e-8)


def test_ragged_and_empty_shapes():
    """Edge cases: a single point, sizes around the 64/128 tile edges, D not a multiple of 4, odd leading dimensions."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels
    g = torch.Generator().manual_seed(5)
    for n1, n2, d in ((1, 1, 1), (1, 130, 3), (63, 129, 7), (64, 128, 16), (65, 127, 17), (200, 1, 33)):
        X = torch.rand(n1, d, generator=g, dtype=torch.float64)
        X2 = torch.rand(n2, d, generator=g, dtype=torch.float64)
        kern = kernels.Matern52(d, ARD=True, length_scales=0.5 + np.arange(d) * 0.05, variance=0.7)
        K = kern.K(X.cuda(), X2.cuda()).detach().cpu()
        ref = O.cov("Matern52", X, X2, torch.as_tensor(0.5 + np.arange(d) * 0.05), torch.tensor([0.7], dtype=torch.float64))
        assert K.shape == (n1, n2) and rel_err(K.numpy(), ref.numpy()) < 1e-13
    # float32 inputs are promoted to float64 (SURVEY 10 "dtype promotion")
    kern = kernels.Rbf(2)
    K = kern.K(torch.rand(4, 2, dtype=torch.float64).cuda(), torch.rand(5, 2, dtype=torch.float32).cuda())
    assert K.dtype == torch.float64
