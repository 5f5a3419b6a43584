"""not-gpu: the C-ABI shared library loads and exports every symbol include/mcb200.h
declares (no compute calls - there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_loads_and_exports_header_symbols():
    from mcircuit_b200 import build, runtime
    path = build.build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "libmcb200.so does not export %s" % n
    assert sorted(runtime.EXPORTED_SYMBOLS) == names
    lib.mcb_version.restype = ctypes.c_int
    assert lib.mcb_version() == 1


def test_library_carries_sm100a_code_only():
    import shutil
    import subprocess
    from mcircuit_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching CUDA."""
    from mcircuit_b200 import runtime
    lib = runtime.load_library()
    st = runtime.McbState()
    st.n_words, st.n_persist, st.n_slots = 4, 2, 1          # n_slots < n_persist
    assert lib.mcb_eval_levels(None, None, 0, 0, ctypes.byref(st), None) == -1
    assert b"geometry" in lib.mcb_last_error()
    assert lib.mcb_set_tuning(b"no_such_knob", 1) == -1
    old = lib.mcb_set_tuning(b"level_unroll", 4)
    assert lib.mcb_set_tuning(b"level_unroll", old) == 4


def test_product_has_no_cpu_fallback():
    """Without a GPU the engine must refuse to construct (no silent CPU path)."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mcircuit_b200 import circuits
    from mcircuit_b200.engine import B200Engine
    from mcircuit_b200.runtime import EngineUnavailable
    with pytest.raises(EngineUnavailable):
        B200Engine(circuits.counter_circuit(), 1, True)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under mcircuit_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "mcircuit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libmcref" not in text and "_cpu_runtime" not in text, f
