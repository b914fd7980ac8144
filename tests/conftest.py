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
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def wfb():
    """The CUDA path. Never falls back: a missing library or device is an error in a gpu test."""
    import torch
    from windflow_b200 import build, ops
    build.build()
    assert torch.cuda.is_available(), "gpu test without a CUDA device"
    return ops
