import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference, when /root/reference is present (never on the GPU box)."""
    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference not present")
    return refshim.load()


@pytest.fixture(scope="session")
def cuda_rt():
    from mcircuit_b200.runtime import CudaRuntime
    return CudaRuntime()          # raises loudly if the .so or the GPU is missing
