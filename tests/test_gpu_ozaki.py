"""The EXPERIMENTAL int8-sliced engine (csrc/gpb_ozaki.cu; off by default): FP64-equivalent products on the INT8 tensor
path.  Checked against the FP64 DMMA engine / torch fp64 and -- at the model level -- against the same reference pins
and tolerances as the default path (loss 1e-9, gradients 1e-7)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture()
def nv():
    from gptorch_b200 import _native
    yield _native
    _native.ozaki_config(0)          # never leak the experimental engine into other tests


def _cov_like(m, k, g, shift=0.0):
    X = torch.rand(max(m, k), 8, generator=g, dtype=torch.float64)
    return torch.exp(-0.5 * torch.cdist(X[:m] + shift, X[:k]).pow(2)).cuda()


def test_sliced_gemm_matches_dgemm(nv):
    g = torch.Generator().manual_seed(21)
    m, n, k = 1024, 512, 256
    A = _cov_like(m, k, g) * torch.exp(2 * torch.randn(m, 1, generator=g, dtype=torch.float64)).cuda()   # row scales differ
    B = _cov_like(n, k, g, 0.1)
    ref = A @ B.t()
    top = float(ref.abs().max())
    err = {s: float((nv.gemm_ozaki_nt(A, B, slices=s) - ref).abs().max()) / top for s in (3, 5, 8)}
    assert err[8] < 5e-14 and err[5] < 1e-8 and err[3] > err[5] > err[8]          # 7 bits per slice
    # alpha / beta and an existing C
    C0 = torch.randn(m, n, generator=g, dtype=torch.float64).cuda()
    C1 = nv.gemm_ozaki_nt(A, B, slices=8, alpha=-1.0, beta=0.5, C=C0.clone())
    assert float((C1 - (0.5 * C0 - ref)).abs().max()) / top < 5e-14
    # A A^T, lower 128-blocks only: the rest of C is not touched
    Cl = torch.full((m, m), 7.0, dtype=torch.float64, device="cuda")
    nv.gemm_ozaki_nt(A, None, slices=8, C=Cl, lower_only=True)
    full = A @ A.t()
    blk = torch.arange(m, device="cuda") // 128
    mask = blk[None, :] <= blk[:, None]
    assert float(((Cl - full).abs() * mask).max()) / float(full.abs().max()) < 5e-14
    assert bool((Cl[~mask] == 7.0).all())
    # non-finite input poisons its row of C instead of being laundered into numbers
    A2 = A.clone()
    A2[5, 7] = float("nan")
    C2 = nv.gemm_ozaki_nt(A2, B, slices=8)
    assert bool(torch.isnan(C2[5]).all()) and bool(torch.isfinite(C2[6]).all())


def test_sliced_gemm_declines_shapes_it_does_not_take(nv):
    from gptorch_b200._lib import NativeLibraryError
    A = torch.rand(64, 100, dtype=torch.float64, device="cuda")              # k % 16 != 0
    with pytest.raises(NativeLibraryError, match="rc=-5"):
        nv.gemm_ozaki_nt(A, None, slices=8)
    with pytest.raises(NativeLibraryError, match="rc=-1"):
        nv.gemm_ozaki_nt(torch.rand(64, 64, dtype=torch.float64, device="cuda"), None, slices=1)
    assert nv.ozaki_config(-1) == 0                                          # off by default


def test_cholesky_and_inverse_on_the_sliced_engine_match_the_dmma_engine(nv):
    """n = 8192 crosses every hook's threshold: trailing updates of the look-ahead Cholesky, the three triangular products of
    the blocked inverse (one of them through the transposing split)."""
    n = 8192
    g = torch.Generator().manual_seed(22)
    X = torch.rand(n, 8, generator=g, dtype=torch.float64).cuda()
    ell = torch.ones(8, dtype=torch.float64, device="cuda")
    s2 = torch.ones(1, dtype=torch.float64, device="cuda")
    noise = torch.full((1,), 0.01, dtype=torch.float64, device="cuda")
    res = {}
    for slices in (0, 8):
        nv.ozaki_config(slices)
        buf, ld = nv._aligned_empty(n, n, X.device)
        nv.kern_fwd(0, X, None, ell, s2, noise=noise, lower=True, out=buf, ldk=ld)
        dinv, info = nv.potrf_(buf, ld)
        assert int(info.item()) == 0
        L = torch.tril(buf[:, :n]).clone()
        kd = nv.potri_(buf, ld, dinv)
        res[slices] = (L, nv.potri_assemble(buf, ld, kd).clone())
    nv.ozaki_config(0)
    (L0, K0), (L8, K8) = res[0], res[8]
    assert float((L8 - L0).abs().max()) / float(L0.abs().max()) < 5e-12
    assert float((K8 - K0).abs().max()) / float(K0.abs().max()) < 1e-11
    assert abs(float(torch.log(L8.diagonal()).sum() - torch.log(L0.diagonal()).sum())) < 1e-9


def test_gpr_reference_pin_holds_on_the_sliced_engine(nv):
    """The N = 8192 reference pin (loss and every gradient of the unmodified reference) at the default tolerances with the
    experimental engine switched on."""
    from conftest import Cases
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import GPR
    n = 8192
    c, nm = Cases("large_cases.npz"), "gpr_n%d" % n
    pin = -6511.334472842767
    X, Y, _ = O.synth_regression(n, 8)
    nv.ozaki_config(8)
    model = GPR(X.numpy(), Y.numpy(), kernels.Rbf(8, ARD=True), likelihood=likelihoods.Gaussian(variance=0.01))
    loss = model.loss()
    loss.backward()
    nv.ozaki_config(0)
    assert abs(loss.item() - pin) <= 1e-9 * abs(pin)
    assert rel_err(model.kernel.variance.grad.cpu().numpy(), c.get(nm, "g_variance")) <= 1e-7
    assert rel_err(model.kernel.length_scales.grad.cpu().numpy(), c.get(nm, "g_length_scales")) <= 1e-7
    assert rel_err(model.likelihood.variance.grad.cpu().numpy(), c.get(nm, "g_noise")) <= 1e-7
