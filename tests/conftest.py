import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref (compiled reference) not built")
    return oracle.ref()


@pytest.fixture(scope="session")
def checker():
    """The compiled reference when oracle/_ref is there (it travels to the GPU box), else the C restatement."""
    import oracle
    return oracle.checker()


@pytest.fixture(scope="session")
def ctx():
    from pycricodecs_b200 import engine
    return engine.Context(0)
