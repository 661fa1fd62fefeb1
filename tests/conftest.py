import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")
    config.addinivalue_line("markers", "slow: full-size configuration (seconds of GPU time)")


@pytest.fixture(scope="session")
def harness():
    import harness as H
    return H
