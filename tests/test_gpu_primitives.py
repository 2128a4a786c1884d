"""functions.cholesky / trtrs / lt_log_determinant / cholesky_inverse and the raw C-ABI primitives on CUDA, checked
against torch CPU (LAPACK) -- i.e. the library the reference dispatches to -- plus full-size checks against the
reference's measured loss pins and size-independent identities."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _spd(n, d=4, noise=0.05, seed=0):
    from oracle import gp_oracle as O
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    K = O.cov("Matern52", X, None, torch.ones(d, dtype=torch.float64), torch.ones(1, dtype=torch.float64))
    return K + noise * torch.eye(n, dtype=torch.float64)


@pytest.mark.parametrize("n", [1, 2, 17, 128, 129, 255, 300, 1000, 1537])
def test_cholesky_solve_logdet_inverse(n):
    from gptorch_b200 import functions
    K = _spd(n)
    Lref = torch.linalg.cholesky(K)
    Kc = K.cuda().requires_grad_(True)
    L = functions.cholesky(Kc)
    assert L.shape == (n, n) and torch.equal(torch.triu(L, 1), torch.zeros_like(L))
    assert rel_err(L.detach().cpu().numpy(), Lref.numpy()) < 1e-12
    g = torch.Generator().manual_seed(1)
    for k in (1, 3, 40):
        b = torch.randn(n, k, generator=g, dtype=torch.float64)
        x = functions.trtrs(b.cuda(), L.detach())
        assert rel_err(x.cpu().numpy(), torch.linalg.solve_triangular(Lref, b, upper=False).numpy()) < 1e-10
        xu = functions.trtrs(b.cuda(), L.detach().t().contiguous(), lower=False)
        assert rel_err(xu.cpu().numpy(), torch.linalg.solve_triangular(Lref.t(), b, upper=True).numpy()) < 1e-10
    ld = functions.lt_log_determinant(L)
    assert ld.ndimension() == 0
    assert ld.item() == pytest.approx(Lref.diag().log().sum().item(), rel=1e-12, abs=1e-13)
    Kinv = functions.cholesky_inverse(L.detach())
    assert rel_err(Kinv.cpu().numpy(), torch.cholesky_inverse(Lref).numpy()) < 1e-9
    Kinv2 = functions.inverse(K.cuda())
    assert rel_err(Kinv2.cpu().numpy(), torch.linalg.inv(K).numpy()) < 1e-9


@pytest.mark.parametrize("n", [5, 200, 700])
def test_cholesky_trtrs_backward(n):
    """Autograd through cholesky + trtrs + log-det vs torch CPU autograd (what the reference relies on)."""
    from gptorch_b200 import functions
    K = _spd(n)
    g = torch.Generator().manual_seed(2)
    for k in (2, 50):
        b = torch.randn(n, k, generator=g, dtype=torch.float64)
        Kc = K.cuda().requires_grad_(True)
        bc = b.cuda().requires_grad_(True)
        L = functions.cholesky(Kc)
        x = functions.trtrs(bc, L)
        (0.5 * x.pow(2).sum() + 3.0 * functions.lt_log_determinant(L)).backward()
        Kr = K.clone().requires_grad_(True)
        br = b.clone().requires_grad_(True)
        Lr = torch.linalg.cholesky(Kr)
        xr = torch.linalg.solve_triangular(Lr, br, upper=False)
        (0.5 * xr.pow(2).sum() + 3.0 * Lr.diag().log().sum()).backward()
        assert rel_err(bc.grad.cpu().numpy(), br.grad.numpy()) < 1e-9
        gk = Kc.grad.cpu()
        gr = 0.5 * (Kr.grad + Kr.grad.t())   # torch returns a symmetrised gradient as well
        assert rel_err((0.5 * (gk + gk.t())).numpy(), gr.numpy()) < 1e-9


def test_not_positive_definite_reports_lapack_info():
    from gptorch_b200 import functions
    A = torch.eye(300, dtype=torch.float64)
    A[200, 200] = -1.0
    with pytest.raises(torch.linalg.LinAlgError, match="order 201"):
        functions._potrf(A.cuda())
    nan = torch.eye(10, dtype=torch.float64)
    nan[3, 3] = float("nan")
    with pytest.raises(RuntimeError):
        functions._potrf(nan.cuda())


@pytest.mark.parametrize("shape", [(128, 128, 16), (300, 200, 100), (77, 130, 1000), (1156, 1280, 256), (2052, 2052, 2048)])
def test_gemm_modes(shape):
    """The DMMA GEMM engine (all three operand layouts, ragged edges, accumulate) vs float64 matmul on the CPU."""
    from gptorch_b200 import _native as nv
    m, n, k = shape
    g = torch.Generator().manual_seed(4)
    A = torch.randn(m, k, generator=g, dtype=torch.float64)
    B = torch.randn(n, k, generator=g, dtype=torch.float64)
    C0 = torch.randn(m, n, generator=g, dtype=torch.float64)
    ref = C0 - A @ B.t()
    for mode, a, b in ((nv.GEMM_NT, A, B), (nv.GEMM_TN, A.t().contiguous(), B.t().contiguous()), (nv.GEMM_NN, A, B.t().contiguous())):
        buf, ld = nv._aligned_empty(m, n, torch.device("cuda"))
        C = buf[:, :n]
        C.copy_(C0)
        nv.gemm(mode, a.cuda(), b.cuda(), alpha=-1.0, beta=1.0, C=C)
        assert rel_err(C.cpu().numpy(), ref.numpy()) < 1e-13
        assert rel_err(nv.gemm(mode, a.cuda(), b.cuda()).cpu().numpy(), (A @ B.t()).numpy()) < 1e-13


def test_gemm_is_race_free_under_repetition():
    """Regression test for the stage-release race (an LDS of a ring stage still in flight when the TMA producer was
    allowed to overwrite it): many repetitions of a ragged, long-k accumulate must stay exact."""
    from gptorch_b200 import _native as nv
    g = torch.Generator().manual_seed(9)
    m = n = 2052
    k = 2048
    A = torch.randn(m, k, generator=g, dtype=torch.float64).cuda()
    B = torch.randn(n, k, generator=g, dtype=torch.float64).cuda()
    C0 = torch.randn(m, n, generator=g, dtype=torch.float64).cuda()
    ref = (C0.cpu() - A.cpu() @ B.cpu().t())
    for _ in range(8):
        buf, ld = nv._aligned_empty(m, n, torch.device("cuda"))
        C = buf[:, :n]
        C.copy_(C0)
        nv.gemm(nv.GEMM_NT, A, B, alpha=-1.0, beta=1.0, C=C)
        assert rel_err(C.cpu().numpy(), ref.numpy()) < 1e-13


@pytest.mark.parametrize("n,pin", [(8192, -6511.334472842767), (16384, -13224.865836863326)])
def test_gpr_loss_matches_reference_pins_at_scale(n, pin):
    """losses the survey measured with the unmodified reference (CPU) on the seeded inputs."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import GPR
    X, Y, _ = O.synth_regression(n, 8)
    model = GPR(X.numpy(), Y.numpy(), kernels.Rbf(8, ARD=True), likelihood=likelihoods.Gaussian(variance=0.01))
    loss = model.loss()
    assert abs(loss.item() - pin) <= 1e-9 * abs(pin)
    loss.backward()
    for p in model.parameters():
        if p.requires_grad:
            assert torch.isfinite(p.grad).all()


def test_full_size_identities_n32768():
    """The named config (N = 32768, D = 8) through size-independent properties: (i) the factor reproduces Ky on a
    random probe vector, (ii) the blocked inverse inverts it, (iii) the analytic gradient agrees with a central
    finite difference of the loss along a random direction of the raw hyper-parameters."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods, _native as nv
    from gptorch_b200.models import GPR
    n, d = 32768, 8
    X, Y, _ = O.synth_regression(n, d)
    Xc = X.cuda()
    ell = torch.ones(d, dtype=torch.float64, device="cuda")
    s2 = torch.ones(1, dtype=torch.float64, device="cuda")
    noise = torch.full((1,), 0.01, dtype=torch.float64, device="cuda")
    buf, ld = nv._aligned_empty(n, n, Xc.device)
    nv.kern_fwd(0, Xc, None, ell, s2, noise=noise, out=buf, ldk=ld)           # full Ky
    v = torch.randn(n, 1, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    Kv = buf[:, :n] @ v                                                       # checker-side matvec
    dinv, info = nv.potrf_(buf, ld)
    assert int(info.item()) == 0
    w = Kv.clone()
    nv.trsv_(buf, dinv, w, False)
    nv.trsv_(buf, dinv, w, True)                                              # Ky^-1 (Ky v) == v
    assert rel_err(w.cpu().numpy(), v.cpu().numpy()) < 1e-9
    kd = nv.potri_(buf, ld, dinv)
    Kinv = nv.potri_assemble(buf, ld, kd)
    assert rel_err((Kinv @ Kv).cpu().numpy(), v.cpu().numpy()) < 1e-8
    del buf, Kinv
    torch.cuda.empty_cache()
    model = GPR(X.numpy(), Y.numpy(), kernels.Rbf(d, ARD=True), likelihood=likelihoods.Gaussian(variance=0.01))
    loss = model.loss()
    loss.backward()
    params = [p for p in model.parameters() if p.requires_grad]
    direction = [torch.randn_like(p) for p in params]
    analytic = sum((p.grad * q).sum().item() for p, q in zip(params, direction))
    eps = 1e-4
    with torch.no_grad():
        for p, q in zip(params, direction):
            p.add_(eps * q)
        lp = model.loss().item()
        for p, q in zip(params, direction):
            p.sub_(2 * eps * q)
        lm = model.loss().item()
    fd = (lp - lm) / (2 * eps)
    assert abs(fd - analytic) <= 1e-5 * abs(analytic)
