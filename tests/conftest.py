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
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def builder():
    """One SvoBuilder context on cuda:0 for the whole GPU session (fails loudly without a B200)."""
    from ooc_svo_builder_b200 import SvoBuilder
    sb = SvoBuilder(0)
    yield sb
    sb.close()
