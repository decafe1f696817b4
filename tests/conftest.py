import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def ref_bces():
    with open(os.path.join(GOLDEN, "ref_bces.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def vignette():
    with open(os.path.join(GOLDEN, "vignette.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def yeast():
    z = np.load(os.path.join(GOLDEN, "yeast_benchmark.npz"))
    return [np.asfortranarray(z[f"y{i:02d}"]) for i in range(1, 18)]


@pytest.fixture(scope="session")
def yeast_names():
    z = np.load(os.path.join(GOLDEN, "yeast_benchmark.npz"))
    return list(z["y01_rownames"]), list(z["y01_colnames"])


@pytest.fixture(scope="session")
def simdata():
    z = np.load(os.path.join(GOLDEN, "simdata.npz"))
    return {k: np.asfortranarray(z[k]) for k in z.files}
