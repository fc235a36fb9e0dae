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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_small():
    with open(os.path.join(GOLDEN, "golden_small.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_readme():
    return np.load(os.path.join(GOLDEN, "golden_readme.npz"))


@pytest.fixture(scope="session")
def golden_random():
    z = np.load(os.path.join(GOLDEN, "golden_random.npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


@pytest.fixture(scope="session")
def oracle():
    from oracle import splitp_oracle
    return splitp_oracle


@pytest.fixture(scope="session")
def golden_rank1():
    with open(os.path.join(GOLDEN, "golden_rank1.json")) as f:
        return json.load(f)
