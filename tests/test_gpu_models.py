"""Parity of the CUDA models against the reference's outputs (tests/golden, made by oracle/make_golden.py).

Tolerances are BASELINE.json's: <= 1e-9 relative on the log marginal likelihood / bound, <= 1e-7 on
hyper-parameter gradients and predictive mean / variance."""
import numpy as np
import pytest
import torch

from conftest import Cases, case_inputs, rel_err

pytestmark = pytest.mark.gpu

LML_TOL = 1e-9
GRAD_TOL = 1e-7
PRED_TOL = 1e-7

_GPR = Cases("gpr_cases.npz")
_VFE = Cases("vfe_cases.npz")
_SVGP = Cases("svgp_cases.npz")


def _kernel(kind, d, ell, var):
    from gptorch_b200 import kernels
    return getattr(kernels, kind)(d, ARD=True, length_scales=np.array(ell, dtype=np.float64).copy(), variance=float(var))


def _grads(model):
    return {n: p.grad.detach().cpu().numpy() for n, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("name", _GPR.names)
def test_gpr_loss_grad_predict(name):
    from gptorch_b200 import likelihoods
    from gptorch_b200.models import GPR
    c = _GPR
    X, Y, g = case_inputs(c, name)
    dy = int(c.get(name, "dy"))
    if dy > 1 and not c.has(name, "Y"):
        pytest.skip("multi-output case without stored Y")
    kind, d = str(c.get(name, "kind")), int(c.get(name, "d"))
    model = GPR(X.numpy(), Y.numpy(), _kernel(kind, d, c.get(name, "ell"), c.get(name, "variance")),
                likelihood=likelihoods.Gaussian(variance=float(c.get(name, "noise"))))
    loss = model.loss()
    assert loss.is_cuda and loss.ndimension() == 1          # test/test_models/test_gpr.py:42
    loss.backward()
    gr = _grads(model)
    lo, g_var, g_ell, g_noise = (loss.detach().cpu().numpy(), gr["kernel.variance"], gr["kernel.length_scales"],
                                 gr["likelihood.variance"])
    if kind != "Exp":
        assert rel_err(lo, c.get(name, "loss")) <= LML_TOL
        assert rel_err(g_var, c.get(name, "g_variance")) <= GRAD_TOL
        assert rel_err(g_ell, c.get(name, "g_length_scales")) <= GRAD_TOL
        assert rel_err(g_noise, c.get(name, "g_noise")) <= GRAD_TOL
    else:
        # Exp/Matern12: the ONE documented deviation.  The reference's K(X) diagonal is
        # sigma2 * exp(-sqrt(round-off of |x|^2 + |x|^2 - 2 x.x)) ~ sigma2 (1 - 1e-8): noise that depends on the BLAS
        # summation order (the reference's CPU and CUDA paths differ from each other by 2.6e-9 .. 5.9e-9 on the LML,
        # profiles/r02_reference_box.json); the CUDA kernel uses the exact r = 0 there.  Proven, not assumed:
        # (1) against the oracle with that single change (O.exact_diagonal) the north-star tolerances hold,
        # (2) against the reference's golden the distance is bounded by the reference's own distance from (1).
        from oracle import gp_oracle as O
        with O.exact_diagonal():
            e_loss, e_gr = O.gpr_loss_and_grads(kind, X, Y, c.get(name, "ell"), float(c.get(name, "variance")),
                                                float(c.get(name, "noise")))
        assert rel_err(lo, e_loss.numpy()) <= LML_TOL
        assert rel_err(g_var, e_gr["variance"].numpy()) <= GRAD_TOL
        assert rel_err(g_ell, e_gr["length_scales"].numpy()) <= GRAD_TOL
        assert rel_err(g_noise, e_gr["noise"].numpy()) <= GRAD_TOL
        assert rel_err(lo, c.get(name, "loss")) <= LML_TOL + rel_err(e_loss.numpy(), c.get(name, "loss"))
        assert rel_err(g_ell, c.get(name, "g_length_scales")) <= GRAD_TOL + rel_err(e_gr["length_scales"].numpy(), c.get(name, "g_length_scales"))
        assert rel_err(g_noise, c.get(name, "g_noise")) <= GRAD_TOL + rel_err(e_gr["noise"].numpy(), c.get(name, "g_noise"))
        if X.shape[0] <= 128:
            # (3) a 40-digit mpmath evaluation (oracle/exact_witness.py): the CUDA result is within tolerance of the
            # exact value while the reference is NOT -- the exception is the reference's noise floor, not ours.
            from oracle import exact_witness as W
            w_loss, w_gr = W.gpr_loss_and_grads(kind, c.get(name, "X"), c.get(name, "Y"), c.get(name, "ell"),
                                                float(c.get(name, "variance")), float(c.get(name, "noise")))
            assert rel_err(lo, w_loss) <= LML_TOL
            assert rel_err(g_var, w_gr["variance"]) <= GRAD_TOL
            assert rel_err(g_ell, w_gr["length_scales"]) <= GRAD_TOL
            assert rel_err(g_noise, w_gr["noise"]) <= GRAD_TOL
            assert rel_err(c.get(name, "loss"), w_loss) > LML_TOL > rel_err(lo, w_loss)
    # loss(x=, y=) equals loss() (test/test_models/test_gpr.py:45-47)
    assert model.loss(x=model.X, y=model.Y).item() == pytest.approx(loss.item(), rel=1e-13)
    if c.has(name, "Xs"):
        Xs = torch.as_tensor(c.get(name, "Xs"))
        with torch.no_grad():
            mu, var = model._predict(Xs.cuda(), diag=True)
            mu2, cov = model._predict(Xs.cuda(), diag=False)
        assert mu.shape == (Xs.shape[0], Y.shape[1]) and var.shape == mu.shape and cov.shape == (Xs.shape[0],) * 2
        assert rel_err(mu.cpu().numpy(), c.get(name, "pred_mean")) <= PRED_TOL
        scale = np.abs(c.get(name, "pred_cov")).max()
        assert np.abs(var.cpu().numpy() - c.get(name, "pred_var")).max() <= PRED_TOL * scale
        assert np.abs(cov.cpu().numpy() - c.get(name, "pred_cov")).max() <= PRED_TOL * scale
        assert rel_err(mu2.cpu().numpy(), c.get(name, "pred_mean")) <= PRED_TOL


def test_gpr_api_contract():
    from gptorch_b200 import kernels, likelihoods, mean_functions
    from gptorch_b200.models import GPR
    rng = np.random.RandomState(0)
    x, y = rng.rand(20, 2), rng.rand(20, 3)
    # numpy, tensors, arbitrary nn.Module mean (test/test_models/test_gpr.py:24-34)
    GPR(x, y, kernels.Rbf(2))
    GPR(torch.as_tensor(x), torch.as_tensor(y), kernels.Rbf(2))
    model = GPR(x, y, kernels.Matern32(2), mean_function=torch.nn.Linear(2, 3).double().cuda())
    loss = model.loss()
    loss.backward()
    assert model.mean_function.weight.grad is not None
    with pytest.raises(ValueError):
        model.loss(x=model.X[:10], y=model.Y)
    # float32 test inputs are promoted (test/test_models/test_gpr.py:61-74)
    model = GPR(x, y, kernels.Matern32(2))
    xs = torch.rand(5, 2, dtype=torch.float32).cuda()
    mu, var = model._predict(xs, diag=True)
    assert mu.dtype == torch.float64 and mu.shape == (5, 3) and var.shape == (5, 3)
    mu, cov = model._predict(xs, diag=False)
    assert cov.shape == (5, 5)
    # predict_* follow the caller's type (test/test_models/test_base.py:83-107)
    out = model.predict_f(rng.rand(4, 2))
    assert isinstance(out[0], np.ndarray) and out[0].shape == (4, 3)
    out = model.predict_y(torch.rand(4, 2, dtype=torch.float64))
    assert out[0].device.type == "cpu"
    s = model.predict_f_samples(rng.rand(4, 2), n_samples=6)
    assert s.shape == (6, 4, 3)
    s = model.predict_y_samples(rng.rand(4, 2), n_samples=2)
    assert s.shape == (2, 4, 3)


def test_gpr_composite_kernel_path():
    """Sum / Product / Linear / Constant kernels use the step-by-step path on native primitives."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import GPR
    X, Y, _ = O.synth_regression(200, 3)
    kern = kernels.Linear(3) + kernels.Rbf(3) + kernels.Constant(3)     # examples/regression_1d.py:42
    model = GPR(X.numpy(), Y.numpy(), kern, likelihood=likelihoods.Gaussian(variance=0.01))
    loss = model.loss()
    loss.backward()
    # oracle with the same composite
    one = torch.ones(1, dtype=torch.float64)
    raw = [torch.zeros(3, dtype=torch.float64, requires_grad=True), torch.zeros(1, dtype=torch.float64, requires_grad=True),
           torch.zeros(1, dtype=torch.float64, requires_grad=True), torch.zeros(1, dtype=torch.float64, requires_grad=True),
           torch.log(torch.tensor([0.01], dtype=torch.float64)).requires_grad_(True)]
    v_lin, var_rbf, ell_rbf, var_c, noise = [r.exp() for r in raw]
    K = O.cov("Linear", X, None, None, v_lin) + O.cov("Rbf", X, None, ell_rbf, var_rbf) + var_c.expand(200, 200)
    L = O.chol(K + noise * torch.eye(200, dtype=torch.float64))
    alpha = O.tri_solve(Y, L)
    ref = 0.5 * alpha.pow(2).sum() + O.tri_logdet(L) + 0.5 * 200 * np.log(2 * np.pi)
    ref.backward()
    assert rel_err(loss.item(), ref.item()) <= LML_TOL
    gr = _grads(model)
    assert rel_err(gr["kernel.kern1.kern1.variance"], raw[0].grad.numpy()) <= GRAD_TOL
    assert rel_err(gr["kernel.kern1.kern2.variance"], raw[1].grad.numpy()) <= GRAD_TOL
    assert rel_err(gr["kernel.kern1.kern2.length_scales"], raw[2].grad.numpy()) <= GRAD_TOL
    assert rel_err(gr["kernel.kern2.variance"], raw[3].grad.numpy()) <= GRAD_TOL
    assert rel_err(gr["likelihood.variance"], raw[4].grad.numpy()) <= GRAD_TOL


@pytest.mark.parametrize("name", _VFE.names)
def test_vfe_loss_grad_predict(name):
    from oracle import gp_oracle as O
    from gptorch_b200 import likelihoods
    from gptorch_b200.models import VFE
    c = _VFE
    X, Y, g = case_inputs(c, name)
    m = int(c.get(name, "m"))
    Z = torch.as_tensor(c.get(name, "Z")) if c.has(name, "Z") else O.synth_inducing(X, m, g)
    kind, d = str(c.get(name, "kind")), int(c.get(name, "d"))
    model = VFE(X.numpy(), Y.numpy(), _kernel(kind, d, c.get(name, "ell"), c.get(name, "variance")),
                inducing_points=Z.numpy(), likelihood=likelihoods.Gaussian(variance=float(c.get(name, "noise"))))
    loss = model.loss()
    assert loss.ndimension() == 0                            # test/test_models/test_sparse_gpr.py:99
    loss.backward()
    gr = _grads(model)
    assert rel_err(loss.item(), c.get(name, "loss")) <= LML_TOL
    assert rel_err(gr["kernel.variance"], c.get(name, "g_variance")) <= GRAD_TOL
    assert rel_err(gr["kernel.length_scales"], c.get(name, "g_length_scales")) <= GRAD_TOL
    assert rel_err(gr["likelihood.variance"], c.get(name, "g_noise")) <= GRAD_TOL
    assert rel_err(gr["Z"], c.get(name, "g_Z")) <= GRAD_TOL
    Xs = torch.as_tensor(c.get(name, "Xs"))
    with torch.no_grad():
        mu, var = model._predict(Xs.cuda(), diag=True)
        _, cov = model._predict(Xs.cuda(), diag=False)
    scale = np.abs(c.get(name, "pred_cov")).max()
    assert rel_err(mu.cpu().numpy(), c.get(name, "pred_mean")) <= PRED_TOL
    assert np.abs(var.cpu().numpy() - c.get(name, "pred_var")).max() <= PRED_TOL * scale
    assert np.abs(cov.cpu().numpy() - c.get(name, "pred_cov")).max() <= PRED_TOL * scale


def test_vfe_reference_known_answer(fixtures):
    """The reference's own tiny VFE case: loss pin and predictions (test/test_models/test_sparse_gpr.py:80-142)."""
    from gptorch_b200 import kernels, likelihoods, mean_functions
    from gptorch_b200.models import VFE
    x = fixtures.raw("sparse/x")[:, None]; y = fixtures.raw("sparse/y")[:, None]; z = fixtures.raw("sparse/z")[:, None]
    kern = kernels.Matern32(1)
    model = VFE(x, y, kern, inducing_points=z, likelihood=likelihoods.Gaussian(variance=1.0), mean_function=mean_functions.Zero(1))
    loss = model.loss()
    assert loss.ndimension() == 0
    assert loss.item() == pytest.approx(8.842242323920674)                 # the reference's pin (rel 1e-6)
    assert loss.item() == pytest.approx(float(fixtures.raw("run/vfe_loss")), rel=1e-12)   # what the reference returns today
    loss_xy = model.loss(x=model.X, y=model.Y)
    assert loss_xy.item() == loss.item()
    with pytest.raises(ValueError):
        model.loss(x=model.X[:1])
    xs = torch.as_tensor(fixtures.raw("sparse/x_test")[:, None]).cuda()
    mu, var = model._predict(xs, diag=True)
    mu2, cov = model._predict(xs, diag=False)
    exp_mu, exp_s = fixtures.raw("sparse/vfe_y_mean"), fixtures.raw("sparse/vfe_y_cov").reshape(2, 2)
    assert mu.detach().cpu().numpy().ravel() == pytest.approx(exp_mu.ravel())
    assert var.detach().cpu().numpy().ravel() == pytest.approx(np.diag(exp_s))
    assert cov.detach().cpu().numpy() == pytest.approx(exp_s)


@pytest.mark.parametrize("name", _SVGP.names)
def test_svgp_loss_grad_predict(name):
    from gptorch_b200 import likelihoods
    from gptorch_b200.models import SVGP
    c = _SVGP
    X, Y = c.get(name, "X"), c.get(name, "Y")
    kind, d = str(c.get(name, "kind")), int(c.get(name, "d"))
    np.random.seed(0)
    model = SVGP(X, Y, _kernel(kind, d, c.get(name, "ell"), c.get(name, "variance")), inducing_points=c.get(name, "Z"),
                 likelihood=likelihoods.Gaussian(variance=float(c.get(name, "noise"))))
    # same random initial posterior as the reference (np.random.seed(0) stream)
    assert rel_err(model.induced_output_mean.detach().cpu().numpy(), c.get(name, "q_mu")) <= 1e-7
    model.induced_output_mean.data = torch.as_tensor(c.get(name, "q_mu")).cuda()
    model.induced_output_chol_cov.data = torch.as_tensor(c.get(name, "q_sqrt_raw")).cuda()
    xb, yb = torch.as_tensor(c.get(name, "xb")).cuda(), torch.as_tensor(c.get(name, "yb")).cuda()
    loss = model.loss(xb, yb)
    assert loss.ndimension() == 0
    loss.backward()
    gr = _grads(model)
    assert rel_err(loss.item(), c.get(name, "loss")) <= LML_TOL
    assert rel_err(gr["kernel.variance"], c.get(name, "g_variance")) <= GRAD_TOL
    assert rel_err(gr["kernel.length_scales"], c.get(name, "g_length_scales")) <= GRAD_TOL
    assert rel_err(gr["likelihood.variance"], c.get(name, "g_noise")) <= GRAD_TOL
    assert rel_err(gr["Z"], c.get(name, "g_Z")) <= GRAD_TOL
    assert rel_err(gr["induced_output_mean"], c.get(name, "g_q_mu")) <= GRAD_TOL
    assert rel_err(gr["induced_output_chol_cov"], c.get(name, "g_q_sqrt_raw")) <= GRAD_TOL
    Xs = torch.as_tensor(c.get(name, "Xs")).cuda()
    with torch.no_grad():
        mu, var = model._predict(Xs, diag=True)
        _, cov = model._predict(Xs, diag=False)
    scale = np.abs(c.get(name, "pred_cov")).max()
    assert rel_err(mu.cpu().numpy(), c.get(name, "pred_mean")) <= PRED_TOL
    assert np.abs(var.cpu().numpy() - c.get(name, "pred_var")).max() <= PRED_TOL * scale
    assert np.abs(cov.cpu().numpy() - c.get(name, "pred_cov")).max() <= PRED_TOL * scale


def test_svgp_reference_known_answer(fixtures):
    """test/test_models/test_sparse_gpr.py:190-292: loss pin 9.5346..., minibatch == full, predictions."""
    from gptorch_b200 import kernels, likelihoods, mean_functions
    from gptorch_b200.models import SVGP
    x = fixtures.raw("sparse/x")[:, None]; y = fixtures.raw("sparse/y")[:, None]; z = fixtures.raw("sparse/z")[:, None]
    model = SVGP(x, y, kernels.Matern32(1), inducing_points=z, likelihood=likelihoods.Gaussian(variance=1.0),
                 mean_function=mean_functions.Zero(1))
    model.induced_output_mean.data = torch.as_tensor(fixtures.raw("sparse/q_mu")[:, None]).cuda()
    l_s = torch.as_tensor(fixtures.raw("sparse/l_s").reshape(2, 2)).cuda()
    model.induced_output_chol_cov.data = model.induced_output_chol_cov._transform.inv(l_s)
    loss = model.loss()
    assert loss.ndimension() == 0
    assert loss.item() == pytest.approx(9.534628739243518)
    assert loss.item() == pytest.approx(float(fixtures.raw("run/svgp_loss")), rel=1e-12)
    model.batch_size = 3   # a minibatch of the full size gives the full loss
    assert model.loss().item() == pytest.approx(loss.item(), rel=1e-12)
    xs = torch.as_tensor(fixtures.raw("sparse/x_test")[:, None]).cuda()
    mu, var = model._predict(xs, diag=True)
    _, cov = model._predict(xs, diag=False)
    exp_mu, exp_s = fixtures.raw("sparse/svgp_y_mean"), fixtures.raw("sparse/svgp_y_cov").reshape(2, 2)
    assert mu.detach().cpu().numpy().ravel() == pytest.approx(exp_mu.ravel())
    assert var.detach().cpu().numpy().ravel() == pytest.approx(np.diag(exp_s))
    assert cov.detach().cpu().numpy() == pytest.approx(exp_s)


def _check_large_predictions(model, c, nm, d):
    """model._predict on the seeded test points of oracle/make_golden_large.py against the unmodified reference."""
    g = torch.Generator().manual_seed(777)
    Xs = torch.rand(48, d, generator=g, dtype=torch.float64).cuda()
    with torch.no_grad():
        mu, var = model._predict(Xs, diag=True)
        _, cov = model._predict(Xs, diag=False)
    scale = np.abs(c.get(nm, "pred_cov")).max()
    assert rel_err(mu.cpu().numpy(), c.get(nm, "pred_mean")) <= PRED_TOL
    assert np.abs(var.cpu().numpy() - c.get(nm, "pred_var")).max() <= PRED_TOL * scale
    assert np.abs(cov.cpu().numpy() - c.get(nm, "pred_cov")).max() <= PRED_TOL * scale


def test_vfe_named_size_pin_n100000_m1024():
    """BASELINE.md section 3's VFE pin (N = 1e5, D = 16, M = 1024 -> 383333.82272224966) and every gradient of the
    unmodified reference at that size (oracle/make_golden_large.py)."""
    from conftest import Cases
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import VFE
    c, nm = Cases("large_cases.npz"), "vfe_n100000_m1024"
    assert abs(c.get(nm, "loss").item() - 383333.82272224966) <= 1e-12 * 383333.8
    X, Y, g = O.synth_regression(100000, 16)
    Z = O.synth_inducing(X, 1024, g)
    model = VFE(X.numpy(), Y.numpy(), kernels.Rbf(16, ARD=True), inducing_points=Z.numpy(),
                likelihood=likelihoods.Gaussian(variance=0.01))
    loss = model.loss()
    loss.backward()
    gr = _grads(model)
    assert rel_err(loss.item(), 383333.82272224966) <= LML_TOL
    assert rel_err(gr["kernel.variance"], c.get(nm, "g_variance")) <= GRAD_TOL
    assert rel_err(gr["kernel.length_scales"], c.get(nm, "g_length_scales")) <= GRAD_TOL
    assert rel_err(gr["likelihood.variance"], c.get(nm, "g_noise")) <= GRAD_TOL
    assert rel_err(gr["Z"], c.get(nm, "g_Z")) <= GRAD_TOL


@pytest.mark.parametrize("name", _VFE.names + ["vfe_n100000_m1024"])
def test_vfe_phi_form_meets_the_parity_tolerances(name):
    """settings.vfe_phi_form = True (Phi = Kuf Kfu streamed, one M x M congruence; 3 N M^2 flop instead of the reference
    order's 5 N M^2): loss <= 1e-9 and every gradient <= 1e-7 against the unmodified reference on every VFE golden,
    including BASELINE.md's N = 1e5 / M = 1024 pin."""
    from conftest import Cases
    from oracle import gp_oracle as O
    from gptorch_b200 import likelihoods, settings
    from gptorch_b200.models import VFE
    large = name.startswith("vfe_n100000")
    c = Cases("large_cases.npz") if large else _VFE
    if large:
        X, Y, g = O.synth_regression(100000, 16)
        Z, kind, d, ell, var, noise = O.synth_inducing(X, 1024, g), "Rbf", 16, np.ones(16), 1.0, 0.01
    else:
        X, Y, g = case_inputs(c, name)
        m = int(c.get(name, "m"))
        Z = torch.as_tensor(c.get(name, "Z")) if c.has(name, "Z") else O.synth_inducing(X, m, g)
        kind, d, ell, var, noise = (str(c.get(name, "kind")), int(c.get(name, "d")), c.get(name, "ell"),
                                    float(c.get(name, "variance")), float(c.get(name, "noise")))
    model = VFE(X.numpy(), Y.numpy(), _kernel(kind, d, ell, var), inducing_points=Z.numpy(),
                likelihood=likelihoods.Gaussian(variance=noise))
    settings.vfe_phi_form = True
    try:
        loss = model.loss()
        loss.backward()
    finally:
        settings.vfe_phi_form = "auto"
    gr = _grads(model)
    assert rel_err(loss.item(), c.get(name, "loss")) <= LML_TOL
    assert rel_err(gr["kernel.variance"], c.get(name, "g_variance")) <= GRAD_TOL
    assert rel_err(gr["kernel.length_scales"], c.get(name, "g_length_scales")) <= GRAD_TOL
    assert rel_err(gr["likelihood.variance"], c.get(name, "g_noise")) <= GRAD_TOL
    assert rel_err(gr["Z"], c.get(name, "g_Z")) <= GRAD_TOL


@pytest.mark.parametrize("chunk", [1000, 999, 64])
def test_vfe_streaming_over_many_ragged_chunks_and_outputs(chunk, monkeypatch):
    """The streamed statistics (gpb_kuf_stats_fwd / _bwd and the reference-order loop) over MANY row chunks with a ragged
    last chunk, several outputs and with / without the panel cache: both forms against the oracle on the same data."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods, settings, _autograd as ag
    from gptorch_b200.models import VFE, sparse_gpr
    n, d, m, dy = 4321, 5, 37, 3
    X, Y1, g = O.synth_regression(n, d)
    Y = torch.cat([Y1 * (1.0 + 0.5 * k) + 0.05 * k for k in range(dy)], dim=1)
    Z = O.synth_inducing(X, m, g)
    ell = 0.5 + 0.1 * np.arange(d)
    h = O.Hyper("Matern52", ell, 1.3, 0.05)
    Zp = Z.clone().requires_grad_(True)
    ref = -O.vfe_elbo(h, X, Y, Zp)
    ref.backward()
    monkeypatch.setattr(sparse_gpr, "VFE_CHUNK_ROWS", chunk)
    for form in (True, False):
        for budget in (ag.VFE_PANEL_CACHE_BYTES, 0):
            monkeypatch.setattr(ag, "VFE_PANEL_CACHE_BYTES", budget)
            model = VFE(X.numpy(), Y.numpy(), kernels.Matern52(d, ARD=True, length_scales=ell.copy(), variance=1.3),
                        inducing_points=Z.numpy(), likelihood=likelihoods.Gaussian(variance=0.05))
            settings.vfe_phi_form = form
            try:
                loss = model.loss()
                loss.backward()
            finally:
                settings.vfe_phi_form = "auto"
            assert rel_err(loss.item(), ref.item()) <= LML_TOL, (form, budget)
            assert rel_err(model.Z.grad.cpu().numpy(), Zp.grad.numpy()) <= GRAD_TOL, (form, budget)
            assert rel_err(model.kernel.length_scales.grad.cpu().numpy(), h.raw_ell.grad.numpy()) <= GRAD_TOL, (form, budget)
            assert rel_err(model.kernel.variance.grad.cpu().numpy(), h.raw_var.grad.numpy()) <= GRAD_TOL, (form, budget)
            assert rel_err(model.likelihood.variance.grad.cpu().numpy(), h.raw_noise.grad.numpy()) <= GRAD_TOL, (form, budget)


def test_vfe_phi_form_is_gated_by_the_conditioning_of_kuu():
    """"auto": the Phi form is used only when the estimated cond_2(Kuu) is below settings.vfe_phi_cond_max.  The
    estimate (two power iterations on the native matvec) is checked against the exact condition number; an
    ill-conditioned Kuu (long length scales) takes the reference order and still matches the oracle."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods, settings, _autograd as ag, _native as nv
    from gptorch_b200.models import VFE
    n, d, m = 6000, 4, 64
    X, Y, g = O.synth_regression(n, d)
    Z = O.synth_inducing(X, m, g)
    used = []
    for ell, expect_phi in ((0.4, True), (1.0, False)):
        model = VFE(X.numpy(), Y.numpy(), kernels.Rbf(d, ARD=True, length_scales=ell * np.ones(d)), inducing_points=Z.numpy(),
                    likelihood=likelihoods.Gaussian(variance=0.01))
        assert n >= 16 * m
        timer = nv.PhaseTimer()
        nv.install_timer(timer)
        try:
            loss = model.loss()
        finally:
            nv.install_timer(None)
        phases = timer.totals_ms()
        used.append("vfe_congruence" in phases)
        assert used[-1] == expect_phi and ("vfe_trsm" in phases) == (not expect_phi)
        Kuu = O.cov("Rbf", Z, None, ell * torch.ones(d, dtype=torch.float64), torch.ones(1, dtype=torch.float64))
        ev = torch.linalg.eigvalsh(Kuu)
        exact = float(ev[-1] / ev[0])
        if exact < 1e12:
            assert 0.3 * exact <= model.last_kuu_condition <= 1.05 * exact
        assert (model.last_kuu_condition <= settings.vfe_phi_cond_max) == expect_phi
        h = O.Hyper("Rbf", ell * np.ones(d), 1.0, 0.01)
        ref = -O.vfe_elbo(h, X, Y, Z)
        assert rel_err(loss.item(), ref.item()) <= LML_TOL
    assert used == [True, False]


def test_svgp_named_size_pin_m2048_b16384():
    """SVGP Matern52-ARD at configs[3]'s M = 2048, D = 32 on a minibatch of 16384: loss and every gradient of the
    unmodified reference; the M x M gradient of the raw Cholesky factor through its diagonal, 8 seeded projections
    and its Frobenius norm."""
    from conftest import Cases
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import SVGP
    c, nm = Cases("large_cases.npz"), "svgp_m2048_b16384"
    n, d, m, batch = 65536, 32, 2048, 16384
    X, Y, g = O.synth_regression(n, d)
    Z = O.synth_inducing(X, m, g)
    idx = torch.randperm(n, generator=g)[:batch]
    np.random.seed(0)
    model = SVGP(X.numpy(), Y.numpy(), kernels.Matern52(d, ARD=True), inducing_points=Z.numpy(),
                 likelihood=likelihoods.Gaussian(variance=0.01), batch_size=batch)
    q_mu, raw = O.seeded_q(m, 1)
    model.induced_output_mean.data.copy_(q_mu)
    model.induced_output_chol_cov.data.copy_(raw)
    loss = model.loss(X[idx].cuda(), Y[idx].cuda())
    loss.backward()
    gr = _grads(model)
    assert rel_err(loss.item(), c.get(nm, "loss")) <= LML_TOL
    assert rel_err(gr["kernel.variance"], c.get(nm, "g_variance")) <= GRAD_TOL
    assert rel_err(gr["kernel.length_scales"], c.get(nm, "g_length_scales")) <= GRAD_TOL
    assert rel_err(gr["likelihood.variance"], c.get(nm, "g_noise")) <= GRAD_TOL
    assert rel_err(gr["Z"], c.get(nm, "g_Z")) <= GRAD_TOL
    assert rel_err(gr["induced_output_mean"], c.get(nm, "g_q_mu")) <= GRAD_TOL
    G = gr["induced_output_chol_cov"]
    assert rel_err(np.diag(G), c.get(nm, "g_q_sqrt_raw_diag")) <= GRAD_TOL
    assert rel_err(O.projections(G), c.get(nm, "g_q_sqrt_raw_proj")) <= GRAD_TOL
    assert abs(np.linalg.norm(G) - c.get(nm, "g_q_sqrt_raw_fro").item()) <= GRAD_TOL * c.get(nm, "g_q_sqrt_raw_fro").item()
    _check_large_predictions(model, c, nm, d)


@pytest.mark.parametrize("name", _SVGP.names + ["svgp_m2048_b16384"])
def test_svgp_quadratic_form_meets_the_parity_tolerances(name):
    """settings.svgp_quadratic_form = True (C = Kuu^-1 (S - Kuu) Kuu^-1 first, one panel product; 3 B M^2 flop instead of
    the reference order's 6 B M^2): loss <= 1e-9 and every gradient <= 1e-7 against the unmodified reference on every
    SVGP golden, including the M = 2048 / D = 32 pin."""
    from conftest import Cases
    from oracle import gp_oracle as O
    from gptorch_b200 import likelihoods, kernels, settings
    from gptorch_b200.models import SVGP
    large = name.startswith("svgp_m2048")
    np.random.seed(0)
    if large:
        c = Cases("large_cases.npz")
        n, d, m, batch = 65536, 32, 2048, 16384
        X, Y, g = O.synth_regression(n, d)
        Z = O.synth_inducing(X, m, g)
        idx = torch.randperm(n, generator=g)[:batch]
        model = SVGP(X.numpy(), Y.numpy(), kernels.Matern52(d, ARD=True), inducing_points=Z.numpy(),
                     likelihood=likelihoods.Gaussian(variance=0.01), batch_size=batch)
        q_mu, raw = O.seeded_q(m, 1)
        xb, yb = X[idx].cuda(), Y[idx].cuda()
    else:
        c = _SVGP
        kind, d = str(c.get(name, "kind")), int(c.get(name, "d"))
        model = SVGP(c.get(name, "X"), c.get(name, "Y"), _kernel(kind, d, c.get(name, "ell"), c.get(name, "variance")),
                     inducing_points=c.get(name, "Z"), likelihood=likelihoods.Gaussian(variance=float(c.get(name, "noise"))))
        q_mu, raw = torch.as_tensor(c.get(name, "q_mu")), torch.as_tensor(c.get(name, "q_sqrt_raw"))
        xb, yb = torch.as_tensor(c.get(name, "xb")).cuda(), torch.as_tensor(c.get(name, "yb")).cuda()
    model.induced_output_mean.data.copy_(q_mu)
    model.induced_output_chol_cov.data.copy_(raw)
    settings.svgp_quadratic_form = True
    try:
        loss = model.loss(xb, yb)
        loss.backward()
    finally:
        settings.svgp_quadratic_form = "auto"
    gr = _grads(model)
    assert rel_err(loss.item(), c.get(name, "loss")) <= LML_TOL
    assert rel_err(gr["kernel.variance"], c.get(name, "g_variance")) <= GRAD_TOL
    assert rel_err(gr["kernel.length_scales"], c.get(name, "g_length_scales")) <= GRAD_TOL
    assert rel_err(gr["likelihood.variance"], c.get(name, "g_noise")) <= GRAD_TOL
    assert rel_err(gr["Z"], c.get(name, "g_Z")) <= GRAD_TOL
    assert rel_err(gr["induced_output_mean"], c.get(name, "g_q_mu")) <= GRAD_TOL
    G = gr["induced_output_chol_cov"]
    if large:
        assert rel_err(np.diag(G), c.get(name, "g_q_sqrt_raw_diag")) <= GRAD_TOL
        assert rel_err(O.projections(G), c.get(name, "g_q_sqrt_raw_proj")) <= GRAD_TOL
    else:
        assert rel_err(G, c.get(name, "g_q_sqrt_raw")) <= GRAD_TOL


def test_svgp_bound_with_a_custom_likelihood_uses_propagate_log():
    """Only the package's Gaussian has expected_log_density; any other Likelihood subclass goes through the abstract
    propagate_log(Normal(mean, sqrt(var)), y) exactly as the reference calls it (gptorch/models/sparse_gpr.py:274-281)."""
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import SVGP

    class Wrapped(likelihoods.Likelihood):
        def __init__(self):
            super().__init__()
            self.inner = likelihoods.Gaussian(variance=0.3)
            self.calls = 0

        def predict_mean_variance(self, mean_f, var_f):
            return self.inner.predict_mean_variance(mean_f, var_f)

        def propagate_log(self, qf, targets):
            self.calls += 1
            assert isinstance(qf, torch.distributions.Normal)
            return self.inner.propagate_log(qf, targets)

    rng = np.random.RandomState(3)
    x, y, z = rng.rand(120, 2), rng.rand(120, 2), rng.rand(9, 2)
    np.random.seed(0)
    a = SVGP(x, y, kernels.Matern32(2), inducing_points=z, likelihood=likelihoods.Gaussian(variance=0.3))
    np.random.seed(0)
    lik = Wrapped()
    b = SVGP(x, y, kernels.Matern32(2), inducing_points=z, likelihood=lik)
    b.induced_output_mean.data.copy_(a.induced_output_mean.data)
    b.induced_output_chol_cov.data.copy_(a.induced_output_chol_cov.data)
    la, lb = a.loss(), b.loss()
    assert lik.calls == 2                                   # one per output dimension
    assert lb.item() == pytest.approx(la.item(), rel=1e-12)
    lb.backward()
    assert lik.inner.variance.grad is not None


def test_jitter_schedule_matches_reference():
    """functions.cholesky: un-jittered try, then +1e-10 ... ; -I never succeeds (SURVEY 10 'Jitter')."""
    from gptorch_b200 import functions
    ones = torch.ones(4, 4, dtype=torch.float64).cuda()
    L = functions.cholesky(ones)
    ref = torch.linalg.cholesky(torch.ones(4, 4, dtype=torch.float64) + 1e-10 * torch.eye(4, dtype=torch.float64))
    assert rel_err(L.cpu().numpy(), ref.numpy()) <= 1e-6   # ill-conditioned by construction (pivots ~1e-10)
    assert torch.equal(torch.triu(L, 1), torch.zeros_like(L))
    with pytest.raises(RuntimeError, match="Max tries exceeded."):
        functions.cholesky(-torch.eye(5, dtype=torch.float64).cuda())
    with pytest.raises(RuntimeError):
        functions._potrf(ones)


def test_optimize_runs():
    """optimize(max_iter=2) with a torch optimiser and with scipy L-BFGS-B (test/test_models/test_base.py:50-53)."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels
    from gptorch_b200.models import GPR, VFE
    X, Y, _ = O.synth_regression(60, 2)
    model = GPR(X.numpy(), Y.numpy(), kernels.Rbf(2))
    l0 = model.loss().item()
    model.optimize(method="Adam", max_iter=2, verbose=False)
    res = model.optimize(method="L-BFGS-B", max_iter=5, verbose=False)
    assert res.fun < l0
    vfe = VFE(X.numpy(), Y.numpy(), kernels.Rbf(2), num_inducing_points=8)
    vfe.optimize(method="L-BFGS-B", max_iter=3, verbose=False)


@pytest.mark.parametrize("model_type,initial,final", [("GPR", 4969.7905, -69.3117), ("VFE", 4969.7974, -69.3109)])
def test_example_regression_1d_reaches_the_reference_optimum(model_type, initial, final):
    """BASELINE config #1: the reference's example (N = 100, Linear + Rbf + Constant, L-BFGS-B) goes
    4969.7905 -> -69.3117 (GPR) and 4969.7974 -> -69.3109 (VFE) [BASELINE.md section 2, measured]."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("regression_1d", os.path.join(ROOT, "examples", "regression_1d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.run(model_type)
    assert out["initial_loss"] == pytest.approx(initial, rel=1e-7)
    assert out["final_loss"] == pytest.approx(final, abs=2e-3)
    assert out["mu"].shape == (200, 1) and out["var"].shape == (200, 1) and np.all(out["var"] > 0)
    assert out["samples"].shape == (5, 200, 1)


@pytest.mark.parametrize("family", ["GPR", "VFE", "SVGP"])
def test_predict_factor_cache(family):
    """under no_grad the parameter-dependent factorisations of _predict are computed once per
    parameter/data state -- the second call launches fewer kernels and returns identical numbers; changing a
    hyper-parameter (even through .data) or the targets invalidates the cache; with autograd on nothing is cached."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods, _native as nv
    from gptorch_b200.models import GPR, VFE, SVGP
    X, Y, g = O.synth_regression(700, 3)
    Z = O.synth_inducing(X, 40, g).numpy()
    np.random.seed(0)

    def build():
        kern = kernels.Matern52(3, ARD=True, length_scales=np.array([0.6, 0.9, 1.2]), variance=1.4)
        lik = likelihoods.Gaussian(variance=0.05)
        if family == "GPR":
            return GPR(X.numpy(), Y.numpy(), kern, likelihood=lik)
        if family == "VFE":
            return VFE(X.numpy(), Y.numpy(), kern, inducing_points=Z, likelihood=lik)
        return SVGP(X.numpy(), Y.numpy(), kern, inducing_points=Z, likelihood=lik)

    model = build()
    xs = torch.rand(50, 3, dtype=torch.float64, generator=g).cuda()
    with torch.no_grad():
        nv.reset_launch_count()
        m1, v1 = model._predict(xs, diag=True)
        first = nv.launch_count()
        nv.reset_launch_count()
        m2, v2 = model._predict(xs, diag=True)
        second = nv.launch_count()
        assert second < first
        assert torch.equal(m1, m2) and torch.equal(v1, v2)
        # a parameter edit through .data (no version bump) must invalidate
        model.kernel.variance.data += 0.3
        m3, _ = model._predict(xs, diag=True)
        fresh = build()
        fresh.kernel.variance.data += 0.3
        if family == "SVGP":            # q(u) is initialised from random points: share it
            fresh.induced_output_mean.data.copy_(model.induced_output_mean.data)
            fresh.induced_output_chol_cov.data.copy_(model.induced_output_chol_cov.data)
        m4, _ = fresh._predict(xs, diag=True)
        assert not torch.equal(m3, m1)
        assert rel_err(m3.cpu().numpy(), m4.cpu().numpy()) < 1e-12
        if family != "SVGP":            # SVGP's posterior does not read Y at prediction time
            model.Y.mul_(2.0)
            m5, _ = model._predict(xs, diag=True)
            assert rel_err(m5.cpu().numpy(), 2.0 * m3.cpu().numpy()) < 1e-9
    # autograd on: recomputed, differentiable
    mu, _ = model._predict(xs, diag=True)
    mu.sum().backward()
    assert model.kernel.length_scales.grad is not None


def _dist_gpr_worker(rank, world, port, n, panel, ret):
    import os
    import torch.distributed as dist
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import GPR, DistributedGPR
    torch.cuda.set_device(rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        X, Y, _ = O.synth_regression(n, 4)
        mk = lambda: kernels.Matern52(4, ARD=True, length_scales=np.array([0.5, 0.8, 1.1, 1.4]), variance=1.3)  # noqa: E731
        lik = lambda: likelihoods.Gaussian(variance=0.02)  # noqa: E731
        dm = DistributedGPR(X.numpy(), Y.numpy(), mk(), likelihood=lik(), panel=panel)
        loss = dm.loss()
        loss.sum().backward()
        sm = GPR(X.numpy(), Y.numpy(), mk(), likelihood=lik())
        ref = sm.loss()
        ref.sum().backward()
        errs = [abs(loss.item() - ref.item()) / abs(ref.item())]
        compared = 0
        for a, b in zip(dm.parameters(), sm.parameters()):
            if b.grad is not None:
                errs.append(float((a.grad - b.grad).abs().max() / b.grad.abs().max()))
                compared += 1
        assert compared == 3        # kernel variance, length scales, noise
        with torch.no_grad():                       # loss-only evaluation frees the slabs without a backward pass
            again = dm.loss().item()
            # distributed prediction against the single-GPU posterior (several ragged blocks of test points)
            Xs = torch.rand(37, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(3)).cuda()
            mu_ref, var_ref = sm._predict(Xs, diag=True)
            _, cov_ref = sm._predict(Xs, diag=False)
            mu, var = dm._predict(Xs, diag=True, block=16)
            mu2, cov = dm._predict(Xs, diag=False)
            scale = float(cov_ref.abs().max())
            pred = [float((mu - mu_ref).abs().max() / mu_ref.abs().max()), float((mu2 - mu_ref).abs().max() / mu_ref.abs().max()),
                    float((var - var_ref).abs().max()) / scale, float((cov - cov_ref).abs().max()) / scale]
            assert max(pred) <= PRED_TOL, pred
            out = dm.predict_y(Xs.cpu().numpy())        # public wrapper: numpy in, numpy out
            assert out[0].shape == (37, 1) and np.all(out[1] > 0)
        errs.append(abs(again - loss.item()) / abs(loss.item()))
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,panel", [(700, 256), (1100, 128), (300, 384)])
def test_distributed_gpr_loss_and_gradient(n, panel):
    """Block-column-cyclic Cholesky -> inverse -> gradient (gptorch_b200/models/dist_gpr.py) on the native library
    over NCCL: every visible GPU is a rank (a single GPU still runs the whole panel pipeline with itself as the only
    owner); loss and all hyper-parameter gradients must match the single-GPU fused node."""
    import socket
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)       # every visible GPU is a rank (8 on the full box)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_dist_gpr_worker, args=(world, port, n, panel, ret), nprocs=world, join=True)
    for rank in range(world):
        errs = ret[rank]
        assert errs[0] <= LML_TOL and errs[-1] <= 1e-13
        assert max(errs[1:-1]) <= GRAD_TOL


# ----------------------------------------------------------------------------------------------------------
# optimiser bridge: CUDA-graph replay of loss() + backward()  (gptorch_b200/model.py GraphedEvaluation)
# ----------------------------------------------------------------------------------------------------------
def _bridge_model(family, n=160):
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import GPR, VFE
    X, Y, g = O.synth_regression(n, 2)
    if family == "GPR-composite":
        return GPR(X.numpy(), Y.numpy(), kernels.Linear(2) + kernels.Rbf(2) + kernels.Constant(2))
    if family == "GPR":
        return GPR(X.numpy(), Y.numpy(), kernels.Matern32(2, ARD=True), likelihood=likelihoods.Gaussian(variance=0.05))
    return VFE(X.numpy(), Y.numpy(), kernels.Rbf(2, ARD=True), inducing_points=O.synth_inducing(X, 12, g).numpy(),
               likelihood=likelihoods.Gaussian(variance=0.05))


@pytest.mark.parametrize("family", ["GPR-composite", "GPR", "VFE"])
def test_graphed_optimizer_bridge_matches_eager(family, capsys):
    """Model._loss_and_grad replays a captured CUDA graph; values and gradients must equal the eager evaluation bit
    for bit (same kernels, same order), for every parameter vector, and the Params must hold the new values."""
    from gptorch_b200 import settings
    rng = np.random.RandomState(3)
    graphed, eager = _bridge_model(family), _bridge_model(family)
    theta0 = graphed._get_param_array()
    assert graphed._graphable()
    for k in range(4):
        theta = theta0 + 0.2 * k * rng.randn(theta0.size)
        settings.cuda_graphs = True
        try:
            f1, g1 = graphed._loss_and_grad(theta)
        finally:
            settings.cuda_graphs = False
        try:
            f2, g2 = eager._loss_and_grad(theta)
        finally:
            settings.cuda_graphs = True
        assert "_graph_eval" in graphed.__dict__ and "_graph_eval" not in eager.__dict__
        assert f1 == f2 and np.array_equal(g1, g2)
        assert np.array_equal(graphed._get_param_array(), theta)
    capsys.readouterr()


def test_graphed_bridge_falls_back_for_jitter_and_uncapturable_models(capsys):
    """(1) Ky not positive-definite without jitter: the replay reports the failed factorisation and the evaluation is
    redone eagerly with the reference's jitter schedule.  (2) A model whose loss() reads the device from the host
    cannot be captured: a warning, then eager evaluation."""
    import warnings
    from gptorch_b200 import kernels, likelihoods, settings
    from gptorch_b200.models import GPR
    x = np.linspace(0, 1, 40).reshape(-1, 1)
    x = np.concatenate([x, x])                         # duplicated points: K is singular
    y = np.sin(6 * x)
    mk = lambda: GPR(x, y, kernels.Rbf(1), likelihood=likelihoods.Gaussian(variance=1e-30))  # noqa: E731
    a, b = mk(), mk()
    theta = a._get_param_array()
    f1, g1 = a._loss_and_grad(theta)
    settings.cuda_graphs = False
    try:
        f2, g2 = b._loss_and_grad(theta)
    finally:
        settings.cuda_graphs = True
    assert np.isfinite(f1) and f1 == f2 and np.array_equal(g1, g2)

    class HostRead(torch.nn.Module):
        def forward(self, x):
            return torch.zeros(x.shape[0], 1, dtype=torch.float64, device=x.device) + float(x.sum().item()) * 0.0

    m = GPR(x[:40], y[:40], kernels.Rbf(1), mean_function=HostRead(), likelihood=likelihoods.Gaussian(variance=0.1))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        f, g = m._loss_and_grad(m._get_param_array())
    assert any("CUDA-graph capture" in str(x.message) for x in w)
    assert m.__dict__.get("_graph_state") == "failed" and np.isfinite(f) and np.all(np.isfinite(g))
    f_again, _ = m._loss_and_grad(m._get_param_array())      # stays on the eager path, still works
    assert f_again == f
    # the GPU is still healthy for other models after the failed capture
    ok = _bridge_model("GPR")
    fo, go = ok._loss_and_grad(ok._get_param_array())
    assert np.isfinite(fo) and "_graph_eval" in ok.__dict__
    capsys.readouterr()


# ----------------------------------------------------------------------------------------------------------
# edge cases: empty / ragged / wide inputs
# ----------------------------------------------------------------------------------------------------------
def test_empty_and_single_point_inputs():
    """Zero test points, one training point, zero-row K: shapes follow the reference (torch broadcasting rules)."""
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import GPR
    rng = np.random.RandomState(0)
    x, y = rng.rand(30, 3), rng.rand(30, 2)
    model = GPR(x, y, kernels.Matern52(3, ARD=True), likelihood=likelihoods.Gaussian(variance=0.1))
    empty = torch.zeros(0, 3, dtype=torch.float64).cuda()
    with torch.no_grad():
        mu, var = model._predict(empty, diag=True)
        assert mu.shape == (0, 2) and var.shape == (0, 2)
        mu, cov = model._predict(empty, diag=False)
        assert mu.shape == (0, 2) and cov.shape == (0, 0)
        assert model.kernel.K(empty).shape == (0, 0)
        assert model.kernel.K(empty, model.X).shape == (0, 30) and model.kernel.K(model.X, empty).shape == (30, 0)
        assert model.kernel.Kdiag(empty).shape == (0,)
    one = GPR(x[:1], y[:1], kernels.Rbf(3), likelihood=likelihoods.Gaussian(variance=0.1))
    loss = one.loss()
    loss.backward()
    # closed form for n = 1: Ky = sigma2 + noise
    ky = 1.0 + 0.1
    want = 0.5 * (y[:1] ** 2).sum() / ky + 0.5 * 2 * np.log(ky) + 0.5 * 2 * np.log(2 * np.pi)
    assert abs(loss.item() - want) < 1e-12 * abs(want)
    with torch.no_grad():
        mu, var = one._predict(torch.as_tensor(x[1:4]).cuda(), diag=True)
    assert mu.shape == (3, 2) and bool((var > 0).all())


@pytest.mark.parametrize("d,dy,n", [(1, 1, 257), (33, 7, 300), (160, 2, 140), (9, 5, 129), (4, 9, 200), (32, 8, 130)])
def test_gpr_wide_inputs_and_many_outputs(d, dy, n):
    """Input dimensions that are not multiples of the staging chunk (and the largest the backward kernel stages,
    D = 160), more outputs than one trsv group (4): loss and gradients against the oracle."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import GPR
    g = torch.Generator().manual_seed(d * 100 + dy)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    Y = torch.randn(n, dy, generator=g, dtype=torch.float64)
    ell = (1.0 + 0.5 * torch.rand(d, generator=g, dtype=torch.float64)).numpy() * np.sqrt(d)
    model = GPR(X.numpy(), Y.numpy(), kernels.Matern32(d, ARD=True, length_scales=ell.copy(), variance=1.3),
                likelihood=likelihoods.Gaussian(variance=0.05))
    loss = model.loss()
    loss.backward()
    h = O.Hyper("Matern32", ell, 1.3, 0.05)
    ref = -O.gpr_loglik(h, X, Y)
    ref.sum().backward()
    assert rel_err(loss.detach().cpu().numpy(), ref.detach().numpy()) <= LML_TOL
    assert rel_err(model.kernel.length_scales.grad.cpu().numpy(), h.raw_ell.grad.numpy()) <= GRAD_TOL
    assert rel_err(model.kernel.variance.grad.cpu().numpy(), h.raw_var.grad.numpy()) <= GRAD_TOL
    assert rel_err(model.likelihood.variance.grad.cpu().numpy(), h.raw_noise.grad.numpy()) <= GRAD_TOL


def test_too_wide_inputs_fail_loudly():
    """D beyond the backward kernel's shared-memory staging (160) is rejected with an error, never silently wrong."""
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200._lib import NativeLibraryError
    from gptorch_b200.models import GPR
    rng = np.random.RandomState(0)
    model = GPR(rng.rand(40, 200), rng.rand(40, 1), kernels.Rbf(200, ARD=True), likelihood=likelihoods.Gaussian(variance=0.1))
    loss = model.loss()                       # forward has no limit on D
    assert np.isfinite(loss.item())
    with pytest.raises(NativeLibraryError):
        loss.backward()


def test_svgp_data_on_host_stream_matches_resident_minibatches():
    """with data_on_host=True the rows stay in pinned host memory and minibatches are gathered and
    copied one step ahead (HostBatchStream); the sequence of minibatches -- hence every loss and gradient -- is the
    same as with device-resident data and the same host RNG stream."""
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import SVGP
    X, Y, g = O.synth_regression(5000, 4)
    Z = O.synth_inducing(X, 24, g).numpy()

    def build(on_host):
        np.random.seed(0)
        return SVGP(X.numpy(), Y.numpy(), kernels.Matern52(4, ARD=True), inducing_points=Z,
                    likelihood=likelihoods.Gaussian(variance=0.05), batch_size=700, data_on_host=on_host)

    def run(model):
        np.random.seed(123)
        out = []
        for _ in range(4):
            for p in model.parameters():
                p.grad = None
            loss = model.loss()
            loss.backward()
            out.append((loss.item(), model.Z.grad.clone(), model.kernel.length_scales.grad.clone()))
        return out

    resident, hosted = build(False), build(True)
    assert resident.X.is_cuda and not hosted.X.is_cuda and hosted.X.is_pinned()
    assert hosted.compute_device.type == "cuda" and hosted.num_data == 5000 and not hosted._graphable()
    a, b = run(resident), run(hosted)
    for (la, gza, gla), (lb, gzb, glb) in zip(a, b):
        assert la == lb and torch.equal(gza, gzb) and torch.equal(gla, glb)
    assert len({l for l, _, _ in a}) == 4                    # four different minibatches
    with torch.no_grad():                                    # prediction never touches the host data
        mu, var = hosted.predict_f(X[:9].numpy())
    assert mu.shape == (9, 1) and np.all(var > 0)
    hosted.batch_size = None
    with pytest.raises(ValueError):
        hosted.loss()


def test_device_kmeans_for_large_data():
    """util.kmeans_centers_device (used for the default inducing inputs above 200k rows): centres land on the clusters,
    are reproducible under the numpy seed (to summation order), and a VFE built without inducing_points uses it."""
    from gptorch_b200 import kernels, util
    from gptorch_b200.models import VFE
    rng = np.random.RandomState(1)
    true = rng.rand(6, 3) * 10.0
    lab = rng.randint(0, 6, size=60000)
    x = true[lab] + 0.05 * rng.randn(60000, 3)
    np.random.seed(5)
    c1 = util.kmeans_centers_device(x, 6, iters=25, chunk=7000)       # ragged chunks
    np.random.seed(5)
    c2 = util.kmeans_centers_device(x, 6, iters=25, chunk=7000)
    assert c1.shape == (6, 3) and np.allclose(c1, c2, rtol=1e-12, atol=1e-12)   # index_add_ sums in atomic order
    # Lloyd's iteration: every centre is (nearly) the mean of the rows nearest to it
    d2 = ((x[:, None, :] - c1[None, :, :]) ** 2).sum(-1)
    a = d2.argmin(1)
    for j in range(6):
        if (a == j).any():
            assert np.allclose(c1[j], x[a == j].mean(0), atol=1e-3)     # converged to (almost) a fixed point
    # inertia no worse than assigning to the true centres by more than the noise level
    assert d2.min(1).mean() < 4 * 3 * 0.05 ** 2 + ((x - true[lab]) ** 2).sum(1).mean() * 50
    old = util.KMEANS_HOST_MAX_ROWS
    import gptorch_b200.models.sparse_gpr as sg
    sg.KMEANS_HOST_MAX_ROWS = 1000
    try:
        np.random.seed(2)
        m = VFE(x[:5000], rng.rand(5000, 1), kernels.Rbf(3), num_inducing_points=6)
    finally:
        sg.KMEANS_HOST_MAX_ROWS = old
    assert m.Z.shape == (6, 3) and m.Z.is_cuda and np.isfinite(m.loss().item())


def test_graphed_bridge_rebuilds_when_storage_is_replaced(capsys):
    """The captured graph reads parameters and data at fixed addresses: replacing a Param's storage or the data
    tensors must rebuild it (never replay against stale memory)."""
    graphed = _bridge_model("GPR")
    theta = graphed._get_param_array()
    f0, g0 = graphed._loss_and_grad(theta)
    first = graphed.__dict__["_graph_eval"]
    # same storage: the same graph object is replayed
    graphed._loss_and_grad(theta + 0.1)
    assert graphed.__dict__["_graph_eval"] is first
    # a Param gets new storage (what p.data = ... does in user code)
    graphed.kernel.variance.data = graphed.kernel.variance.data.clone()
    f1, g1 = graphed._loss_and_grad(theta)
    second = graphed.__dict__["_graph_eval"]
    assert second is not first and f1 == f0 and np.array_equal(g1, g0)
    # new targets: the loss must change accordingly (no replay against the old Y)
    graphed.Y = graphed.Y * 2.0
    f2, _ = graphed._loss_and_grad(theta)
    fresh = _bridge_model("GPR")
    fresh.Y = fresh.Y * 2.0
    f3, _ = fresh._loss_and_grad(theta)
    assert graphed.__dict__["_graph_eval"] is not second and f2 == f3 and f2 != f0
    capsys.readouterr()
