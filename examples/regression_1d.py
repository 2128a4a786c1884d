"""1-D GP regression demo on the B200 package (BASELINE config #1; same data, kernel and optimiser as the
reference's examples/regression_1d.py: N = 100, Linear + Rbf + Constant, L-BFGS-B <= 100 iterations).

    python examples/regression_1d.py [--model-type GPR|VFE] [--plot]
"""
import os
import sys
from argparse import ArgumentParser

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from gptorch_b200 import kernels  # noqa: E402
from gptorch_b200.models import GPR, VFE  # noqa: E402


def f(x):
    return np.sin(2.0 * np.pi * x) + np.cos(3.5 * np.pi * x) - 3.0 * x + 5.0


def run(model_type="GPR", max_iter=100, plot=False):
    torch.manual_seed(42)
    np.random.seed(42)
    n = 100
    x = np.linspace(0, 1, n).reshape((-1, 1))
    y = f(x) + 0.1 * np.random.randn(n, 1)
    kern = kernels.Linear(1) + kernels.Rbf(1) + kernels.Constant(1)
    model = GPR(x, y, kern) if model_type == "GPR" else VFE(x, y, kern)
    initial = model.loss().item()
    result = model.optimize(method="L-BFGS-B", max_iter=max_iter)
    print("Trained model:")
    print(model)
    x_test = np.linspace(-1, 2, 200).reshape((-1, 1))
    with torch.no_grad():
        mu, s = model.predict_y(x_test)
        y_samp = model.predict_y_samples(x_test, n_samples=5)
    if plot:
        import matplotlib.pyplot as plt
        unc = 2.0 * np.sqrt(s)
        xt = x_test.flatten()
        plt.fill_between(xt, (mu - unc).flatten(), (mu + unc).flatten(), color=(0.9,) * 3)
        plt.plot(xt, mu)
        plt.plot(xt, f(xt))
        for ys in y_samp:
            plt.plot(xt, ys, color=(0.4, 0.7, 1.0), alpha=0.5)
        plt.plot(x, y, "o")
        plt.show()
    return {"initial_loss": initial, "final_loss": float(result.fun), "evals": int(result.nfev), "mu": mu, "var": s,
            "samples": y_samp}


if __name__ == "__main__":
    parser = ArgumentParser()
    parser.add_argument("--model-type", type=str, choices=("GPR", "VFE"), default="GPR")
    parser.add_argument("--plot", action="store_true")
    a = parser.parse_args()
    out = run(a.model_type, plot=a.plot)
    print("loss %.4f -> %.4f in %d evaluations" % (out["initial_loss"], out["final_loss"], out["evals"]))
