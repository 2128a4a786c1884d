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


@pytest.mark.parametrize("shape", [(2600, 2500, 300), (3072, 1704, 130), (2560, 2630, 64)])
def test_gemm_half_tile_configuration(shape):
    """Launches of >= 300 tiles run the 128 x 64 half-tile configuration (two CTAs per SM): all three operand layouts,
    ragged edges on both sides of a half boundary (N mod 128 below and above 64), accumulate, vs float64 matmul on the CPU."""
    from gptorch_b200 import _native as nv
    m, n, k = shape
    assert ((m + 127) // 128) * ((n + 127) // 128) >= 300
    g = torch.Generator().manual_seed(21)
    A = torch.randn(m, k, generator=g, dtype=torch.float64)
    B = torch.randn(n, k, generator=g, dtype=torch.float64)
    C0 = torch.randn(m, n, generator=g, dtype=torch.float64)
    ref = 0.5 * C0 - A @ B.t()
    for mode, a, b in ((nv.GEMM_NT, A, B), (nv.GEMM_TN, A.t().contiguous(), B.t().contiguous()), (nv.GEMM_NN, A, B.t().contiguous())):
        buf, ld = nv._aligned_empty(m, n, torch.device("cuda"))
        C = buf[:, :n]
        C.copy_(C0)
        nv.gemm(mode, a.cuda(), b.cuda(), alpha=-1.0, beta=0.5, C=C)
        assert rel_err(C.cpu().numpy(), ref.numpy()) < 1e-13


def test_gemm_half_tile_lower_and_triangular_k_ranges():
    """The same configuration for the SYRK / LAUUM shapes: lower tiles only, and the k-range flags of triangular operands."""
    from gptorch_b200 import _native as nv
    n, k = 3400, 500                      # 27 x 28 / 2 = 378 lower tiles
    g = torch.Generator().manual_seed(22)
    A = torch.randn(n, k, generator=g, dtype=torch.float64)
    C0 = torch.randn(n, n, generator=g, dtype=torch.float64)
    C = nv._aligned_empty(n, n, torch.device("cuda"))[0][:, :n]
    C.copy_(C0)
    nv.gemm(nv.GEMM_NT, A.cuda(), A.cuda(), alpha=-1.0, beta=1.0, C=C, lower_only=True)
    ref = C0 - A @ A.t()
    assert rel_err(torch.tril(C).cpu().numpy(), torch.tril(ref).numpy()) < 1e-13
    # T T^T with T upper triangular (k >= row and k >= column): the LAUUM product of the blocked inverse
    T = torch.triu(torch.randn(n, n, generator=g, dtype=torch.float64))
    out = nv.gemm(nv.GEMM_NT, T.cuda(), T.cuda(), lower_only=True, flags=nv.GF_KLO_M | nv.GF_KLO_N)
    assert rel_err(torch.tril(out).cpu().numpy(), torch.tril(T @ T.t()).numpy()) < 1e-12
    # X T with T upper triangular (k <= column)
    X = torch.randn(2700, n, generator=g, dtype=torch.float64)
    out = nv.gemm(nv.GEMM_NN, X.cuda(), T.cuda(), flags=nv.GF_KHI_N)
    assert rel_err(out.cpu().numpy(), (X @ T).numpy()) < 1e-12


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


@pytest.mark.parametrize("n", [40, 300, 777])
def test_public_functions_are_differentiable_like_the_reference(n):
    """trtrs(lower=False) with many right-hand sides, cholesky_inverse and inverse backpropagate into the triangular
    / SPD argument exactly as torch CPU autograd does for the reference (gptorch/functions.py:50-58, :71-76)."""
    from gptorch_b200 import functions
    K = _spd(n)
    Lref = torch.linalg.cholesky(K)
    g = torch.Generator().manual_seed(5)
    b = torch.randn(n, 48, generator=g, dtype=torch.float64)
    W = torch.randn(n, n, generator=g, dtype=torch.float64)
    # upper-triangular solve, > TRSV_MAX_RHS columns
    U = Lref.t().contiguous().cuda().requires_grad_(True)
    bc = b.cuda().requires_grad_(True)
    x = functions.trtrs(bc, U, lower=False)
    (x * x).sum().backward()
    Ur = Lref.t().contiguous().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    xr = torch.linalg.solve_triangular(Ur, br, upper=True)
    (xr * xr).sum().backward()
    assert rel_err(x.detach().cpu().numpy(), xr.detach().numpy()) < 1e-10
    assert rel_err(bc.grad.cpu().numpy(), br.grad.numpy()) < 1e-9
    assert rel_err(torch.triu(U.grad).cpu().numpy(), torch.triu(Ur.grad).numpy()) < 1e-9
    # cholesky_inverse
    Lc = Lref.cuda().requires_grad_(True)
    (functions.cholesky_inverse(Lc) * W.cuda()).sum().backward()
    Lr = Lref.clone().requires_grad_(True)
    (torch.cholesky_inverse(Lr) * W).sum().backward()
    assert rel_err(torch.tril(Lc.grad).cpu().numpy(), torch.tril(Lr.grad).numpy()) < 1e-8
    # inverse (through the Cholesky node)
    Kc = K.cuda().requires_grad_(True)
    (functions.inverse(Kc) * W.cuda()).sum().backward()
    Kr = K.clone().requires_grad_(True)
    (torch.linalg.inv(Kr) * W).sum().backward()
    gk, gr = Kc.grad.cpu(), Kr.grad
    assert rel_err((gk + gk.t()).numpy(), (gr + gr.t()).numpy()) < 1e-8


_LARGE = None


def _large():
    global _LARGE
    if _LARGE is None:
        from conftest import Cases
        _LARGE = Cases("large_cases.npz")
    return _LARGE


@pytest.mark.parametrize("n", [8192, 16384])
def test_gpr_loss_and_gradients_match_reference_at_scale(n):
    """Loss AND every gradient of the unmodified reference (CPU, oracle/make_golden_large.py) on SURVEY 8d's seeded
    inputs; the losses are also BASELINE.md section 3's pins."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import GPR
    c, nm = _large(), "gpr_n%d" % n
    pin = {8192: -6511.334472842767, 16384: -13224.865836863326}[n]
    assert abs(c.get(nm, "loss").item() - pin) <= 1e-12 * abs(pin)
    X, Y, _ = O.synth_regression(n, 8)
    model = GPR(X.numpy(), Y.numpy(), kernels.Rbf(8, ARD=True), likelihood=likelihoods.Gaussian(variance=0.01))
    loss = model.loss()
    assert abs(loss.item() - pin) <= 1e-9 * abs(pin)
    loss.backward()
    assert rel_err(model.kernel.variance.grad.cpu().numpy(), c.get(nm, "g_variance")) <= 1e-7
    assert rel_err(model.kernel.length_scales.grad.cpu().numpy(), c.get(nm, "g_length_scales")) <= 1e-7
    assert rel_err(model.likelihood.variance.grad.cpu().numpy(), c.get(nm, "g_noise")) <= 1e-7
    if c.has(nm, "pred_mean"):      # GPR._predict at this size against the reference (mean, variance, full covariance)
        g = torch.Generator().manual_seed(777)
        Xs = torch.rand(48, 8, generator=g, dtype=torch.float64).cuda()
        with torch.no_grad():
            mu, var = model._predict(Xs, diag=True)
            _, cov = model._predict(Xs, diag=False)
        scale = np.abs(c.get(nm, "pred_cov")).max()
        assert rel_err(mu.cpu().numpy(), c.get(nm, "pred_mean")) <= 1e-7
        assert np.abs(var.cpu().numpy() - c.get(nm, "pred_var")).max() <= 1e-7 * scale
        assert np.abs(cov.cpu().numpy() - c.get(nm, "pred_cov")).max() <= 1e-7 * scale


def test_gpr_headline_config_matches_the_unmodified_reference_n32768():
    """configs[1] (N = 32768, D = 8): loss and every gradient against the unmodified reference run on the GPU box
    (tools/reference_box.py -> tests/golden/gpr_n32768_reference.json: the host-CPU/MKL path where the host had the
    ~77 GB it needs, and the same reference under model.cuda() -> cuSOLVER/cuBLAS)."""
    import json
    import os
    from conftest import GOLDEN
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import GPR
    with open(os.path.join(GOLDEN, "gpr_n32768_reference.json")) as f:
        pins = json.load(f)
    n = 32768
    X, Y, _ = O.synth_regression(n, 8)
    model = GPR(X.numpy(), Y.numpy(), kernels.Rbf(8, ARD=True), likelihood=likelihoods.Gaussian(variance=0.01))
    loss = model.loss()
    loss.backward()
    ours = {"kernel.variance": model.kernel.variance.grad, "kernel.length_scales": model.kernel.length_scales.grad,
            "likelihood.variance": model.likelihood.variance.grad}
    checked = 0
    for arm in ("cpu", "cuda"):
        ref = pins.get(arm)
        if not ref or "loss" not in ref:
            continue
        assert abs(loss.item() - ref["loss"]) <= 1e-9 * abs(ref["loss"]), arm
        for name, g in ours.items():
            assert rel_err(g.cpu().numpy(), np.array(ref["grads"][name])) <= 1e-7, (arm, name)
        checked += 1
    assert checked >= 1


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


def test_row_reductions_and_panel_vector_products():
    """gpb_rowdot / gpb_gemv_n / gpb_rows_scale_add_outer (the variance and mean epilogues of the predictive equations)
    on ragged, padded and unaligned panels against torch."""
    from gptorch_b200 import _native as nv, _autograd as ag
    g = torch.Generator().manual_seed(9)
    for rows, cols, dy in ((1, 1, 1), (37, 129, 3), (300, 1000, 1), (513, 77, 6)):
        A = torch.randn(rows, cols, generator=g, dtype=torch.float64).cuda()
        B = torch.randn(rows, cols, generator=g, dtype=torch.float64).cuda()
        V = torch.randn(cols, dy, generator=g, dtype=torch.float64).cuda()
        G = torch.randn(rows, dy, generator=g, dtype=torch.float64).cuda()
        s = torch.randn(rows, generator=g, dtype=torch.float64).cuda()
        assert rel_err(nv.rowdot(A, B).cpu().numpy(), (A * B).sum(1).cpu().numpy()) < 1e-13
        pad = torch.zeros(rows, cols + 3, dtype=torch.float64, device="cuda")
        pad[:, 1:cols + 1] = A                                   # unaligned base, padded rows
        assert rel_err(nv.rowdot(pad[:, 1:cols + 1], B).cpu().numpy(), (A * B).sum(1).cpu().numpy()) < 1e-13
        assert rel_err(nv.gemv_n(A, V).cpu().numpy(), (A @ V).cpu().numpy()) < 1e-13
        ref = 2.0 * s[:, None] * A + G @ V.t()
        out = nv.rows_scale_add_outer_(A.clone(), s=s, scale=2.0, G=G, V=V)
        assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 1e-13
        Ac = A.clone().requires_grad_(True)
        (ag.RowSumSqFn.apply(Ac) * s).sum().backward()
        assert rel_err(Ac.grad.cpu().numpy(), (2.0 * s[:, None] * A).cpu().numpy()) < 1e-13


def test_triangular_inverse_and_syrk_nodes_match_autograd():
    """TriInvTFn (T = L^-T) and SyrkFn (U U^T), the M x M building blocks of the sparse models' quadratic forms."""
    from gptorch_b200 import _autograd as ag, _native as nv
    n = 300
    K = _spd(n)
    Lref = torch.linalg.cholesky(K)
    g = torch.Generator().manual_seed(3)
    W = torch.randn(n, n, generator=g, dtype=torch.float64)
    Lc = Lref.cuda().requires_grad_(True)
    T = ag.TriInvTFn.apply(Lc, nv.tri_diag_inverse(nv._c(Lc.detach())))
    C = ag.SyrkFn.apply(T, True)
    ((T * W.cuda()).sum() + (C * W.cuda()).sum()).backward()
    Lr = Lref.clone().requires_grad_(True)
    Tr = torch.linalg.solve_triangular(Lr, torch.eye(n, dtype=torch.float64), upper=False).t()
    ((Tr * W).sum() + ((Tr @ Tr.t()) * W).sum()).backward()
    assert rel_err(T.detach().cpu().numpy(), Tr.detach().numpy()) < 1e-10
    assert rel_err(C.detach().cpu().numpy(), (Tr @ Tr.t()).detach().numpy()) < 1e-10
    assert rel_err(torch.tril(Lc.grad).cpu().numpy(), torch.tril(Lr.grad).numpy()) < 1e-8


def test_distributed_ops_split_long_reductions():
    """NativeOps.gemm (gptorch_b200/models/dist_gpr.py) k-slices TN products whose output has too few tiles for the GPU
    (the early block rows of Ky^-1 = T^T T); the result must equal the plain product."""
    from gptorch_b200.models.dist_gpr import NativeOps
    from gptorch_b200 import _native as nv
    ops = NativeOps()
    g = torch.Generator().manual_seed(12)
    for k, m, n in ((5000, 256, 384), (4096, 130, 77), (9001, 1024, 129)):
        A = torch.randn(k, m, generator=g, dtype=torch.float64).cuda()
        B = torch.randn(k, n, generator=g, dtype=torch.float64).cuda()
        buf = nv._aligned_empty(m, n, A.device)[0]
        C = buf[:, :n]
        nv.reset_launch_count()
        ops.gemm(ops.GEMM_TN, A, B, C=C)
        assert nv.launch_count() == 1                                   # one batched split-K launch
        assert rel_err(C.cpu().numpy(), (A.t() @ B).cpu().numpy()) < 1e-12
    A = torch.randn(300, 256, generator=g, dtype=torch.float64).cuda()      # short reduction: plain product
    B = torch.randn(300, 128, generator=g, dtype=torch.float64).cuda()
    C = nv._aligned_empty(256, 128, A.device)[0][:, :128]
    ops.gemm(ops.GEMM_TN, A, B, C=C)
    assert rel_err(C.cpu().numpy(), (A.t() @ B).cpu().numpy()) < 1e-12
