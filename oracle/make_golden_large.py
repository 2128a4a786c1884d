"""Pins at BASELINE.json's named sizes from the REAL reference (cics-nd/gptorch at /root/reference).

Run in the build container only (the GPU box has no /root/reference; it takes ~3 min and ~16 GB of RAM):

    python oracle/make_golden_large.py            # -> tests/golden/large_cases.npz

Cases (inputs are regenerated from SURVEY 8d's seeds by the tests, only the reference's outputs are stored):
  * gpr_n8192, gpr_n16384     GPR Rbf-ARD D=8, default hyper-parameters: loss + every gradient (the largest sizes
                              the reference runs in this container; N=32768 is made on the GPU box's host by
                              tools/reference_box.py and stored as tests/golden/gpr_n32768_reference.json)
  * vfe_n100000_m1024         BASELINE.md section 3's VFE pin (383333.82272224966): loss + every gradient
  * svgp_m2048_b16384         SVGP Matern52-ARD D=32, M=2048, minibatch of 16384 out of 65536 rows; q(u) is set from
                              a seeded draw so the test can rebuild it.  The M x M gradient is stored through its
                              diagonal, 8 seeded projections and its Frobenius norm.
"""
import os
import sys
import time
import warnings

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore")

import gptorch  # noqa: E402,F401  (the reference)
from gptorch import kernels as rk, likelihoods as rl  # noqa: E402
from gptorch.models import GPR, VFE, SVGP  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402


def grads_of(model):
    return {n: p.grad.detach().numpy().copy() for n, p in model.named_parameters() if p.grad is not None}


seeded_q, projections = O.seeded_q, O.projections   # shared with the tests


def gpr_large(n, out):
    X, Y, _ = O.synth_regression(n, 8)
    model = GPR(X.numpy(), Y.numpy(), rk.Rbf(8, ARD=True), likelihood=rl.Gaussian(variance=0.01))
    t0 = time.perf_counter()
    loss = model.loss()
    loss.backward()
    sec = time.perf_counter() - t0
    gr = grads_of(model)
    nm = "gpr_n%d" % n
    out[nm + "/loss"] = loss.detach().numpy()
    out[nm + "/g_variance"] = gr["kernel.variance"]
    out[nm + "/g_length_scales"] = gr["kernel.length_scales"]
    out[nm + "/g_noise"] = gr["likelihood.variance"]
    out[nm + "/seconds"] = np.array(sec)
    print(nm, repr(loss.item()), "%.1f s" % sec, flush=True)
    if n <= 8192:
        predictions(model, nm, 8, out)


def predictions(model, nm, d, out, n_test=48):
    """model._predict on seeded test points: mean, variance (diag=True) and the full covariance."""
    g = torch.Generator().manual_seed(777)
    Xs = torch.rand(n_test, d, generator=g, dtype=torch.float64)
    with torch.no_grad():
        mu, var = model._predict(Xs, diag=True)
        _, cov = model._predict(Xs, diag=False)
    out[nm + "/pred_mean"], out[nm + "/pred_var"], out[nm + "/pred_cov"] = mu.numpy(), var.numpy().copy(), cov.numpy()


def vfe_large(out, n=100000, d=16, m=1024):
    X, Y, g = O.synth_regression(n, d)
    Z = O.synth_inducing(X, m, g)
    model = VFE(X.numpy(), Y.numpy(), rk.Rbf(d, ARD=True), inducing_points=Z.numpy(), likelihood=rl.Gaussian(variance=0.01))
    t0 = time.perf_counter()
    loss = model.loss()
    loss.backward()
    sec = time.perf_counter() - t0
    gr = grads_of(model)
    nm = "vfe_n%d_m%d" % (n, m)
    out[nm + "/loss"] = loss.detach().numpy()
    out[nm + "/g_variance"] = gr["kernel.variance"]
    out[nm + "/g_length_scales"] = gr["kernel.length_scales"]
    out[nm + "/g_noise"] = gr["likelihood.variance"]
    out[nm + "/g_Z"] = gr["Z"]
    out[nm + "/seconds"] = np.array(sec)
    print(nm, repr(loss.item()), "%.1f s" % sec, flush=True)
    assert abs(loss.item() - 383333.82272224966) <= 1e-9 * 383333.8, "BASELINE.md section 3 pin moved"
    predictions(model, nm, d, out)


def svgp_large(out, n=65536, d=32, m=2048, batch=16384):
    X, Y, g = O.synth_regression(n, d)
    Z = O.synth_inducing(X, m, g)
    idx = torch.randperm(n, generator=g)[:batch]
    xb, yb = X[idx], Y[idx]
    np.random.seed(0)
    model = SVGP(X.numpy(), Y.numpy(), rk.Matern52(d, ARD=True), inducing_points=Z.numpy(),
                 likelihood=rl.Gaussian(variance=0.01), batch_size=batch)
    q_mu, raw = seeded_q(m, 1)
    model.induced_output_mean.data.copy_(q_mu)
    model.induced_output_chol_cov.data.copy_(raw)
    t0 = time.perf_counter()
    loss = model.loss(xb, yb)
    loss.backward()
    sec = time.perf_counter() - t0
    gr = grads_of(model)
    nm = "svgp_m%d_b%d" % (m, batch)
    G = gr["induced_output_chol_cov"]
    out[nm + "/loss"] = loss.detach().numpy()
    out[nm + "/g_variance"] = gr["kernel.variance"]
    out[nm + "/g_length_scales"] = gr["kernel.length_scales"]
    out[nm + "/g_noise"] = gr["likelihood.variance"]
    out[nm + "/g_Z"] = gr["Z"]
    out[nm + "/g_q_mu"] = gr["induced_output_mean"]
    out[nm + "/g_q_sqrt_raw_diag"] = np.diag(G).copy()
    out[nm + "/g_q_sqrt_raw_proj"] = projections(G)
    out[nm + "/g_q_sqrt_raw_fro"] = np.array(np.linalg.norm(G))
    out[nm + "/seconds"] = np.array(sec)
    print(nm, repr(loss.item()), "%.1f s" % sec, flush=True)
    predictions(model, nm, d, out)


def main():
    out = {}
    torch.set_num_threads(os.cpu_count() or 1)
    vfe_large(out)
    svgp_large(out)
    gpr_large(8192, out)
    gpr_large(16384, out)
    # cross-check against the loss pins BASELINE.md section 3 recorded in the survey session
    assert abs(out["gpr_n8192/loss"].item() + 6511.334472842767) <= 1e-9 * 6511.3
    assert abs(out["gpr_n16384/loss"].item() + 13224.865836863326) <= 1e-9 * 13224.9
    out["torch_version"] = np.array(torch.__version__)
    out["threads"] = np.array(torch.get_num_threads())
    np.savez_compressed(os.path.join(OUT, "large_cases.npz"), **out)
    print("large_cases.npz", os.path.getsize(os.path.join(OUT, "large_cases.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
