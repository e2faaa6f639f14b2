import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = os.path.join(HERE, "golden")
    d = {}
    d.update(np.load(os.path.join(g, "st_fixtures.npz")))
    d.update(np.load(os.path.join(g, "oracle_heads.npz")))
    d["images"] = np.load(os.path.join(g, "images_56.npy"))
    return d
