import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    import nexus_b200
    c = nexus_b200.Context(0)   # raises when there is no GPU: the product has no CPU path
    yield c
    c.close()


@pytest.fixture(scope="session")
def have_ref():
    import oracle_lib
    return oracle_lib.have_ref()
