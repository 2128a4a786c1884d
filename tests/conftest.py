import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Cases:
    """Accessor for the flattened "<case>/<field>" npz files written by oracle/make_golden.py."""

    def __init__(self, fname):
        self._z = np.load(os.path.join(GOLDEN, fname), allow_pickle=False)
        self.names = [str(n) for n in self._z["names"]] if "names" in self._z.files else []

    def get(self, case, field, default=None):
        key = "%s/%s" % (case, field)
        return self._z[key] if key in self._z.files else default

    def has(self, case, field):
        return "%s/%s" % (case, field) in self._z.files

    def raw(self, key):
        return self._z[key]


@pytest.fixture(scope="session")
def fixtures():
    return Cases("reference_fixtures.npz")


@pytest.fixture(scope="session")
def gpr_cases():
    return Cases("gpr_cases.npz")


@pytest.fixture(scope="session")
def vfe_cases():
    return Cases("vfe_cases.npz")


@pytest.fixture(scope="session")
def svgp_cases():
    return Cases("svgp_cases.npz")


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def case_inputs(cases, name):
    """X, Y (and the torch generator positioned after them) for a golden case: stored or regenerated from the seed."""
    from oracle import gp_oracle as O
    n, d = int(cases.get(name, "n")), int(cases.get(name, "d"))
    Xr, Yr, g = O.synth_regression(n, d)
    if cases.has(name, "X"):
        X, Y = torch.as_tensor(cases.get(name, "X")), torch.as_tensor(cases.get(name, "Y"))
    else:
        X, Y = Xr, Yr
    return X, Y, g
