"""One-dimensional GP regression on the B200 package -- BASELINE config #0.

The experiment is the reference's examples/regression_1d.py:26-49 (so its numbers can be compared one to one):
100 equally spaced inputs on [0, 1], targets sin(2 pi x) + cos(3.5 pi x) - 3x + 5 plus N(0, 0.1^2) noise drawn after
seeding numpy and torch with 42, covariance Linear + Rbf + Constant, hyper-parameters fitted by L-BFGS-B (at most 100
iterations) on the exact-GP or the VFE objective, then predictions on 200 points of [-1, 2].

    python examples/regression_1d.py [--model-type GPR|VFE] [--plot]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from gptorch_b200 import kernels  # noqa: E402
from gptorch_b200 import models  # noqa: E402

N_TRAIN, N_TEST, NOISE_STD, SEED = 100, 200, 0.1, 42


def truth(x):
    """Noise-free target of the demo."""
    return np.sin(2.0 * np.pi * x) + np.cos(3.5 * np.pi * x) - 3.0 * x + 5.0


def make_data():
    np.random.seed(SEED)
    torch.manual_seed(SEED)
    x = np.linspace(0.0, 1.0, N_TRAIN)[:, None]
    return x, truth(x) + NOISE_STD * np.random.randn(N_TRAIN, 1)


def make_model(kind, x, y):
    covariance = kernels.Linear(1) + kernels.Rbf(1) + kernels.Constant(1)
    return {"GPR": models.GPR, "VFE": models.VFE}[kind](x, y, covariance)


def show(x, y, grid, mean, variance, draws):
    import matplotlib.pyplot as plt
    band = 2.0 * np.sqrt(variance).ravel()
    g, m = grid.ravel(), mean.ravel()
    plt.fill_between(g, m - band, m + band, color="0.9")
    plt.plot(g, m, label="posterior mean")
    plt.plot(g, truth(g), label="truth")
    for d in draws:
        plt.plot(g, d.ravel(), color=(0.4, 0.7, 1.0), alpha=0.5)
    plt.plot(x, y, "o", label="data")
    plt.legend()
    plt.show()


def run(model_type="GPR", max_iter=100, plot=False):
    x, y = make_data()
    model = make_model(model_type, x, y)
    first = model.loss().item()
    fit = model.optimize(method="L-BFGS-B", max_iter=max_iter)
    print("Trained model:")
    print(model)
    grid = np.linspace(-1.0, 2.0, N_TEST)[:, None]
    with torch.no_grad():
        mean, variance = model.predict_y(grid)
        draws = model.predict_y_samples(grid, n_samples=5)
    if plot:
        show(x, y, grid, mean, variance, draws)
    return {"initial_loss": first, "final_loss": float(fit.fun), "evals": int(fit.nfev), "mu": mean, "var": variance,
            "samples": draws}


if __name__ == "__main__":
    cli = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    cli.add_argument("--model-type", choices=("GPR", "VFE"), default="GPR")
    cli.add_argument("--plot", action="store_true")
    opts = cli.parse_args()
    res = run(opts.model_type, plot=opts.plot)
    print("loss %.4f -> %.4f in %d evaluations" % (res["initial_loss"], res["final_loss"], res["evals"]))
