import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from tests import _oracle
    return _oracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference library (oracle/_ref), or None when it was not built / did not travel."""
    from tests import _oracle
    return _oracle.Reference.load_or_none()
