import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/): the checker, never the thing under test."""
    import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def ssb():
    """The product package; the shared library must already be built (python -m soundscope_b200.build)."""
    import soundscope_b200 as S
    S.lib()
    return S


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test running without a CUDA device")
    torch.cuda.set_device(0)
    return torch
