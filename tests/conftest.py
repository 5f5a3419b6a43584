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


def _gpu_runtime_or_reason():
    try:
        from mcircuit_b200.runtime import CudaRuntime
        CudaRuntime()
        return None
    except Exception as exc:           # EngineUnavailable: no B200 here, or libmcb200.so not built
        return "%s: %s" % (type(exc).__name__, exc)


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a B200 skips the gpu-marked tests instead of failing in
    the first one; `-m gpu` on a GPU box runs them, and there the engine itself fails loudly
    (EngineUnavailable) if the extension is missing - no silent fallback either way."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    import torch
    if torch.cuda.is_available():
        return                          # GPU box: run them; a missing .so must FAIL, not skip
    reason = _gpu_runtime_or_reason() or "no CUDA device"
    skip = pytest.mark.skip(reason="needs a B200 (%s)" % reason)
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(autouse=True)
def _shipped_tuning(request):
    """Every gpu test starts AND ends on the shipped kernel tuning (mcb200.cu `Tuning`), so a test
    that sweeps a knob cannot leak it into the tests collected after it."""
    if "gpu" not in request.keywords:
        yield
        return
    from mcircuit_b200 import runtime
    lib = runtime.load_library()
    lib.mcb_reset_tuning()
    yield
    lib.mcb_reset_tuning()
