import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def warm_cpu_oracle():
    """The first multi-threaded torch-CPU evaluation of a process is not reproducible on the GPU box's host
    (oracle.warm_up documents the measurement); take it before any test takes a reference value."""
    import oracle
    oracle.warm_up()
