import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fit-sne_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden_graph():
    import numpy as np
    g = np.load(os.path.join(GOLDEN, "graph_n3000.npz"))
    return g["row"], g["col"], g["val"], g["labels"]


@pytest.fixture(scope="session")
def golden_gradients():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "gradients_n3000.npz"))


@pytest.fixture(scope="session")
def golden_runs():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "runs_n3000.npz"))
