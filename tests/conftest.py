import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def aliked_state():
    from b200slam import weights
    return weights.synthetic_aliked_state("aliked-n16", seed=0)


@pytest.fixture(scope="session")
def lightglue_state():
    from b200slam import weights
    return weights.synthetic_lightglue_state(seed=0)
