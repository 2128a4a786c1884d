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
    # tighter than the reference's allclose for the fused kernels
    if name in ("Matern32", "Matern52", "Rbf", "Linear"):
        assert rel_err(kx, fixtures.raw("kern/%s_kx" % name)) < 1e-13
        assert rel_err(kx2, fixtures.raw("kern/%s_kx2" % name)) < 1e-13
    if name in ("Exp", "Matern12"):
        # everything but the K(X) diagonal is as tight; the diagonal is exactly sigma2 here where the reference carries
        # sigma2 * exp(-sqrt(round-off)) (tests/test_gpu_models.py Exp branch)
        off = ~np.eye(kx.shape[0], dtype=bool)
        assert rel_err(kx[off], fixtures.raw("kern/%s_kx" % name)[off]) < 1e-13
        assert rel_err(kx2, fixtures.raw("kern/%s_kx2" % name)) < 1e-13
        assert np.array_equal(np.diag(kx), np.ones(kx.shape[0]))
        assert np.abs(np.diag(fixtures.raw("kern/%s_kx" % name)) - 1.0).max() < 1e-7


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
    tol = 1e-9
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
    # K(X): the CUDA kernels use the exact r = 0 on the diagonal; O.exact_diagonal() is the oracle with that one change
    # (it only matters for Exp, whose reference diagonal carries sqrt(round-off) ~ 1e-8 noise)
    with O.exact_diagonal():
        (O.cov(name, Xo2, None, raw_ell.exp().detach(), raw_var.exp().detach()) * Gs).sum().backward()
    assert rel_err(Xc2.grad.cpu().numpy(), Xo2.grad.numpy()) < 1e-8


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


# ----------------------------------------------------------------------------------------------------------
# composite kernels: one fused pass (gpb_kern_sop_fwd) and its leaf-by-leaf backward (gpb_kern_bwd_mul)
# ----------------------------------------------------------------------------------------------------------
from conftest import Cases  # noqa: E402

_COMP = Cases("composite_cases.npz")


def _build_composite(c, name):
    from gptorch_b200 import kernels
    kinds = [str(k) for k in c.get(name, "kinds")]
    d = int(c.get(name, "d"))
    leaves = []
    for i, kind in enumerate(kinds):
        var = c.get(name, "leaf%d/variance" % i)
        if kind == "Linear":
            leaves.append(kernels.Linear(d, variance=var.copy(), ARD=True))
        elif kind in ("Constant", "White"):
            leaves.append(getattr(kernels, kind)(d, variance=float(var[0])))
        else:
            leaves.append(getattr(kernels, kind)(d, ARD=True, length_scales=c.get(name, "leaf%d/ell" % i).copy(),
                                                 variance=float(var[0])))
    kern = eval(str(c.get(name, "expr")), {"__builtins__": {}}, {"k%d" % i: k for i, k in enumerate(leaves)})
    return kern, leaves, kinds


def _oracle_composite_exact_diagonal(c, name):
    """K(X), GPR loss and every leaf gradient of a stored composite case from the oracle with the exact K(X) diagonal."""
    from oracle import gp_oracle as O
    kinds = [str(k) for k in c.get(name, "kinds")]
    X, Y = torch.as_tensor(c.get(name, "X")), torch.as_tensor(c.get(name, "Y"))
    n = X.shape[0]
    raws, o_leaves = [], []
    for i, kind in enumerate(kinds):
        r_ell = (torch.log(torch.as_tensor(c.get(name, "leaf%d/ell" % i))).requires_grad_(True)
                 if c.has(name, "leaf%d/ell" % i) else None)
        r_var = torch.log(torch.as_tensor(c.get(name, "leaf%d/variance" % i))).requires_grad_(True)
        raws.append((r_ell, r_var))
        o_leaves.append((kind, None if r_ell is None else r_ell.exp(), r_var.exp()))
    r_noise = torch.log(torch.tensor([float(c.get(name, "noise"))], dtype=torch.float64)).requires_grad_(True)
    with O.exact_diagonal():
        K = O.cov_composite(str(c.get(name, "expr")), o_leaves, X)
        L = O.chol(K + r_noise.exp() * torch.eye(n, dtype=torch.float64))
        alpha = O.tri_solve(Y, L)
        loss = 0.5 * alpha.pow(2).sum() + O.tri_logdet(L) + 0.5 * n * np.log(2 * np.pi)
        loss.backward()
    out = {"Kx": K.detach().numpy(), "loss": loss.detach().numpy().reshape(1), "g_noise": r_noise.grad.numpy()}
    for i, (r_ell, r_var) in enumerate(raws):
        out["leaf%d/g_variance" % i] = r_var.grad.numpy()
        if r_ell is not None:
            out["leaf%d/g_length_scales" % i] = r_ell.grad.numpy()
    return out


@pytest.mark.parametrize("name", _COMP.names)
def test_composite_kernel_forward_and_gpr(name):
    """K(X), K(X, X2), GPR loss, every leaf's hyper-parameter gradient and the GPR prediction of a Sum / Product tree
    against the real reference (tests/golden/composite_cases.npz)."""
    from gptorch_b200 import kernels, likelihoods, _native as nv
    from gptorch_b200.models import GPR
    c = _COMP
    kern, leaves, kinds = _build_composite(c, name)
    assert kernels.sum_of_products(kern) is not None          # takes the fused path
    X, X2 = torch.as_tensor(c.get(name, "X")).cuda(), torch.as_tensor(c.get(name, "X2")).cuda()
    nv.reset_launch_count()
    Kx = kern.K(X)
    assert nv.launch_count() == 1                             # ONE pass for the whole tree
    # Exp leaves: the reference's K(X) diagonal is sigma2 * exp(-sqrt(round-off)) ~ sigma2 (1 - 1e-8); the
    # CUDA kernels use the exact r = 0 there.  For such trees the expectations are recomputed by the oracle with that
    # single change (O.exact_diagonal) and compared at the usual tolerances; against the reference's golden the allowed
    # distance is the tolerance plus the reference's own distance from that variant.
    has_exp = "Exp" in kinds
    exp = _oracle_composite_exact_diagonal(c, name) if has_exp else None

    def close(got, key, tol, scale=None):
        ref = c.get(name, key)
        den = scale if scale is not None else max(np.abs(ref).max(), 1e-300)
        err = lambda a, b: np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max() / den  # noqa: E731
        if exp is None:
            return err(got, ref) < tol
        return err(got, exp[key]) < tol and err(got, ref) < tol + err(exp[key], ref)

    assert close(Kx.detach().cpu().numpy(), "Kx", 1e-12)
    off = ~np.eye(Kx.shape[0], dtype=bool)
    assert rel_err(Kx.detach().cpu().numpy()[off], c.get(name, "Kx")[off]) < 1e-12
    assert rel_err(kern.K(X, X2).detach().cpu().numpy(), c.get(name, "Kx2")) < 1e-12
    model = GPR(c.get(name, "X"), c.get(name, "Y"), kern, likelihood=likelihoods.Gaussian(variance=float(c.get(name, "noise"))))
    loss = model.loss()
    loss.backward()
    assert close(loss.detach().cpu().numpy(), "loss", 1e-9)
    assert close(model.likelihood.variance.grad.cpu().numpy(), "g_noise", 1e-7)
    for i, leaf in enumerate(leaves):
        assert close(leaf.variance.grad.cpu().numpy(), "leaf%d/g_variance" % i, 1e-7), (i, kinds[i])
        if c.has(name, "leaf%d/g_length_scales" % i):
            scale = max(np.abs(c.get(name, "leaf%d/g_length_scales" % i)).max(), np.abs(c.get(name, "g_noise")).max())
            assert close(leaf.length_scales.grad.cpu().numpy(), "leaf%d/g_length_scales" % i, 1e-7, scale), (i, kinds[i])
    with torch.no_grad():
        mu, var = model._predict(torch.as_tensor(c.get(name, "Xs")).cuda(), diag=True)
    assert rel_err(mu.cpu().numpy(), c.get(name, "pred_mean")) < 1e-7
    assert np.abs(var.cpu().numpy() - c.get(name, "pred_var")).max() < 1e-7 * np.abs(c.get(name, "Kx")).max()


@pytest.mark.parametrize("expr,kinds", [("k0 * k1 + k2", ["Matern52", "Periodic", "Linear"]), ("k0 + k1", ["Rbf", "Exp"]),
                                         ("(k0 + k1) * (k2 + k3)", ["Rbf", "Constant", "Matern32", "White"])])
def test_composite_kernel_backward_dense(expr, kinds):
    """Upstream gradient G through K(X, Z) and K(X) of a composite kernel: gradients w.r.t. every leaf parameter and
    w.r.t. X, Z against torch autograd on the oracle's composition (ragged sizes, non-square)."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels
    d, n1, n2 = 5, 203, 77
    g = torch.Generator().manual_seed(7)
    X = torch.rand(n1, d, generator=g, dtype=torch.float64)
    Z = torch.rand(n2, d, generator=g, dtype=torch.float64)
    G = torch.randn(n1, n2, generator=g, dtype=torch.float64)
    Gs = torch.randn(n1, n1, generator=g, dtype=torch.float64)
    vals = []
    for i, kind in enumerate(kinds):
        if kind == "Linear":
            vals.append((kind, None, 0.4 + 0.2 * np.arange(d)))
        elif kind in ("Constant", "White"):
            vals.append((kind, None, np.array([0.5 + 0.1 * i])))
        else:
            vals.append((kind, 0.7 + 0.15 * np.arange(d) + 0.05 * i, np.array([0.9 + 0.3 * i])))
    # oracle
    Xo, Zo = X.clone().requires_grad_(True), Z.clone().requires_grad_(True)
    o_leaves, raws = [], []
    for kind, ell, var in vals:
        e = torch.tensor(ell).requires_grad_(True) if ell is not None else None
        v = torch.tensor(var).requires_grad_(True)
        raws.append((e, v))
        o_leaves.append((kind, e, v))
    # K(X): the oracle with the exact diagonal (O.exact_diagonal) -- only an Exp leaf can tell the difference
    sym = 1.0
    with O.exact_diagonal():
        val = (O.cov_composite(expr, o_leaves, Xo, Zo) * G).sum() + sym * (O.cov_composite(expr, o_leaves, Xo) * Gs).sum()
        val.backward()
    # CUDA
    leaves = []
    for kind, ell, var in vals:
        if kind == "Linear":
            leaves.append(kernels.Linear(d, variance=var.copy(), ARD=True))
        elif kind in ("Constant", "White"):
            leaves.append(getattr(kernels, kind)(d, variance=float(var[0])))
        else:
            leaves.append(getattr(kernels, kind)(d, ARD=True, length_scales=ell.copy(), variance=float(var[0])))
    kern = eval(expr, {"__builtins__": {}}, {"k%d" % i: k for i, k in enumerate(leaves)})
    assert kernels.sum_of_products(kern) is not None
    Xc, Zc = X.cuda().requires_grad_(True), Z.cuda().requires_grad_(True)
    out = (kern.K(Xc, Zc) * G.cuda()).sum() + sym * (kern.K(Xc) * Gs.cuda()).sum()
    assert abs(out.item() - val.item()) <= 1e-11 * abs(val.item()) + 1e-9
    out.backward()
    scale = float(max(Xo.grad.abs().max(), Zo.grad.abs().max()))
    assert float((Xc.grad.cpu() - Xo.grad).abs().max()) < 1e-9 * scale
    assert float((Zc.grad.cpu() - Zo.grad).abs().max()) < 1e-9 * scale
    for leaf, (e, v), (kind, _, _) in zip(leaves, raws, vals):
        # module parameters are raw (log) values: d/d raw = value * d/d value
        gv = leaf.variance.grad.cpu() / leaf.variance.transform().detach().cpu()
        assert float((gv - v.grad).abs().max()) < 1e-9 * max(float(v.grad.abs().max()), scale), kind
        if e is not None:
            ge = leaf.length_scales.grad.cpu() / leaf.length_scales.transform().detach().cpu()
            assert float((ge - e.grad).abs().max()) < 1e-9 * max(float(e.grad.abs().max()), scale), kind


def test_composite_falls_back_for_user_kernels():
    """A leaf that overrides K() is user code: the tree composes tensors like the reference."""
    from gptorch_b200 import kernels

    class Mine(kernels.Rbf):
        def K(self, X, X2=None):
            return 2.0 * super().K(X, X2)

    k = Mine(2) + kernels.Rbf(2)
    assert kernels.sum_of_products(k) is None
    x = torch.rand(9, 2, dtype=torch.float64).cuda()
    assert torch.allclose(k.K(x), 3.0 * kernels.Rbf(2).K(x))


@pytest.mark.parametrize("n1,n2", [(1, 1), (5, 3), (129, 257), (300, 64)])
def test_composite_kernel_ragged_shapes_and_fill(n1, n2):
    """gpb_kern_sop_fwd on ragged / tiny shapes, the lower-only fill with the noise diagonal, and empty inputs."""
    from oracle import gp_oracle as O
    from gptorch_b200 import _native as nv
    d = 4
    g = torch.Generator().manual_seed(n1 * 1000 + n2)
    X = torch.rand(n1, d, generator=g, dtype=torch.float64)
    X2 = torch.rand(n2, d, generator=g, dtype=torch.float64)
    ell = torch.tensor([0.5, 0.8, 1.1, 1.4], dtype=torch.float64)
    one, two = torch.tensor([1.3], dtype=torch.float64), torch.tensor([0.4], dtype=torch.float64)
    v = torch.tensor([0.2, 0.3, 0.4, 0.5], dtype=torch.float64)
    leaves = [("Matern52", ell, one), ("Linear", None, v), ("Periodic", ell[:1], two), ("White", None, two)]
    expr = "k0 * k1 + k2 + k3"
    c = lambda t: None if t is None else t.cuda()  # noqa: E731
    terms = [[(nv.KIND["Matern52"], c(ell), c(one)), (nv.KIND["Linear"], c(v), None)],
             [(nv.KIND["Periodic"], c(ell[:1]), c(two))], [(nv.KIND["White"], None, c(two))]]
    want = O.cov_composite(expr, leaves, X, X2).numpy()
    got = nv.kern_sop_fwd(terms, X.cuda(), X2.cuda()).cpu().numpy()
    assert got.shape == (n1, n2) and rel_err(got, want) < 1e-12
    # symmetric: noise on the diagonal, lower fill leaves the strictly-upper 128-column tiles untouched
    noise = torch.tensor([0.07], dtype=torch.float64)
    want = (O.cov_composite(expr, leaves, X) + 0.07 * torch.eye(n1, dtype=torch.float64)).numpy()
    full = nv.kern_sop_fwd(terms, X.cuda(), None, noise=noise.cuda()).cpu().numpy()
    assert rel_err(full, want) < 1e-12
    buf, ld = nv._aligned_empty(n1, n1, torch.device("cuda"))
    buf.fill_(-7.0)
    low = nv.kern_sop_fwd(terms, X.cuda(), None, noise=noise.cuda(), lower=True, out=buf, ldk=ld).cpu().numpy()
    assert rel_err(np.tril(low), np.tril(want)) < 1e-12
    # empty second argument
    assert nv.kern_sop_fwd(terms, X.cuda(), X2[:0].cuda()).shape == (n1, 0)


def test_composite_kernel_limits():
    """More than 8 terms / 16 leaf evaluations: the tree is composed from tensors instead (same numbers)."""
    from gptorch_b200 import kernels, _native as nv
    x = torch.rand(33, 2, dtype=torch.float64).cuda()
    leaves = [kernels.Rbf(2, variance=0.1 * (i + 1)) for i in range(9)]
    k = leaves[0]
    for leaf in leaves[1:]:
        k = k + leaf
    assert kernels.sum_of_products(k) is None
    ref = sum(leaf.K(x) for leaf in leaves)
    assert torch.allclose(k.K(x), ref, rtol=1e-14, atol=0)
    with pytest.raises(ValueError):
        nv.kern_sop_fwd([[(0, torch.ones(1).double().cuda(), torch.ones(1).double().cuda())]] * 9, x, None)


def test_linear_kdiag_native_forward_and_gradient():
    """Linear.Kdiag (gptorch/kernels.py:264-265) runs gpb_linear_kdiag on CUDA tensors and stays differentiable in
    the variances and the inputs."""
    from gptorch_b200 import kernels, _native as nv
    g = torch.Generator().manual_seed(4)
    X = torch.rand(301, 5, generator=g, dtype=torch.float64)
    w = torch.randn(301, generator=g, dtype=torch.float64)
    var = 0.3 + 0.2 * np.arange(5)
    kern = kernels.Linear(5, variance=var.copy(), ARD=True)
    Xc = X.cuda().requires_grad_(True)
    nv.reset_launch_count()
    kd = kern.Kdiag(Xc)
    assert nv.launch_count() == 1
    (kd * w.cuda()).sum().backward()
    Xr = X.clone().requires_grad_(True)
    raw = torch.log(torch.as_tensor(var)).requires_grad_(True)
    ref = torch.sum(Xr * Xr * raw.exp(), 1)
    (ref * w).sum().backward()
    assert rel_err(kd.detach().cpu().numpy(), ref.detach().numpy()) < 1e-14
    assert rel_err(Xc.grad.cpu().numpy(), Xr.grad.numpy()) < 1e-13
    assert rel_err(kern.variance.grad.cpu().numpy(), raw.grad.numpy()) < 1e-13


@pytest.mark.parametrize("name", ["Rbf", "Matern52"])
def test_forward_square_tile_path(name):
    """Sizes past the point where the D <= 16 forward kernel switches to 128 x 128 tiles (gpb_kern.cu, KF_SQUARE_MIN_TILES):
    a ragged rows x M panel (odd D: scalar row loads) and a symmetric fill, full and lower, against the oracle."""
    from oracle import gp_oracle as O
    from gptorch_b200 import _native as nv
    from gptorch_b200 import kernels
    g = torch.Generator().manual_seed(11)
    kind = {"Rbf": 0, "Matern52": 3}[name]
    # panel: 70001 x 999 (1094 x 8 tiles of 64 x 128 -> square path), D = 5
    d = 5
    ell = torch.as_tensor(0.6 + 0.1 * np.arange(d))
    var = torch.tensor([1.7], dtype=torch.float64)
    X, Z = torch.rand(70001, d, generator=g, dtype=torch.float64), torch.rand(999, d, generator=g, dtype=torch.float64)
    kern = getattr(kernels, name)(d, ARD=True, length_scales=ell.numpy().copy(), variance=1.7)
    K = kern.K(X.cuda(), Z.cuda()).detach().cpu()
    assert K.shape == (70001, 999) and rel_err(K.numpy(), O.cov(name, X, Z, ell, var).numpy()) < 1e-13
    # symmetric: n = 9100 (143 x 72 tiles), D = 8 (vector row loads), noise on the diagonal, full vs lower fill
    n, d = 9100, 8
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    ell = torch.full((d,), 0.8, dtype=torch.float64)
    noise = torch.tensor([0.05], dtype=torch.float64)
    full, ld = nv._aligned_empty(n, n, torch.device("cuda"))
    low, ld2 = nv._aligned_empty(n, n, torch.device("cuda"))
    low.zero_()
    nv.kern_fwd(kind, X.cuda(), None, ell.cuda(), var.cuda(), noise=noise.cuda(), out=full, ldk=ld)
    nv.kern_fwd(kind, X.cuda(), None, ell.cuda(), var.cuda(), noise=noise.cuda(), lower=True, out=low, ldk=ld2)
    with O.exact_diagonal():
        ref = O.cov(name, X, None, ell, var) + noise * torch.eye(n, dtype=torch.float64)
    assert rel_err(full.cpu().numpy(), ref.numpy()) < 1e-13
    assert torch.equal(torch.tril(low), torch.tril(full))                       # the same values, bit for bit
    assert torch.equal(full, full.T) and bool((full.diagonal() == 1.7 + 0.05).all())
    assert float(torch.triu(low, 129).abs().max()) == 0.0                      # nothing beyond the diagonal tiles
    # the same rows from a base address that is 8 but not 16 bytes aligned (scalar row loads): bit-identical output
    for rows in (n, 300):                                                       # square-tile and 64-row-tile paths
        flat = torch.empty(rows * d + 1, dtype=torch.float64, device="cuda")
        Xo = flat[1:].view(rows, d)
        Xo.copy_(X[:rows])
        assert Xo.data_ptr() % 16 == 8 and Xo.is_contiguous()
        Xa = X[:rows].cuda()
        Zc = X[:4000 if rows == n else 257].cuda()                             # 143 x 32 tiles: square path
        Ka = nv.kern_fwd(kind, Xa, Zc, ell.cuda(), var.cuda())
        Ko = nv.kern_fwd(kind, Xo, Zc, ell.cuda(), var.cuda())
        Kz = nv.kern_fwd(kind, Zc, Xo, ell.cuda(), var.cuda())
        assert torch.equal(Ka, Ko) and torch.equal(Kz, Ka.T.contiguous())
