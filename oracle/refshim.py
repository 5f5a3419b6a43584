"""Load the UNMODIFIED reference (matijakevic/mcircuit) from /root/reference under a
compatibility shim.  TEST INFRASTRUCTURE ONLY - used by oracle/gen_golden.py and by the
not-gpu tests that pin oracle/restate.py against the reference's own LLVM JIT when
/root/reference exists (it does not exist on the GPU box; nothing in the product path,
bench.py or the gpu tests imports this module).

The reference targets llvmlite 0.36 (requirements.txt:4).  This image has llvmlite 0.47, where
  * llvmlite.binding.initialize() raises (reference core/simulator.py:11 calls it), and
  * the legacy pass-manager builder used at core/simulator.py:251-256 is gone.
Both are patched at run time; no reference file is edited or copied.  Optimisation level
cannot change results: the emitted IR has no undefined behaviour for in-range pin values.
"""
import contextlib
import importlib
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("MCIRCUIT_REFERENCE", "/root/reference")


def available() -> bool:
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "core", "simulator.py")):
        return False
    try:
        import llvmlite  # noqa: F401
        import networkx  # noqa: F401
    except Exception:
        return False
    return True


def _patch_llvmlite():
    import llvmlite.binding as llvm

    if getattr(llvm, "_mcb200_shimmed", False):
        return
    llvm.initialize = lambda: None

    class _PMB:  # stands in for create_pass_manager_builder (reference simulator.py:251-255)
        inlining_threshold = 0
        opt_level = 0

        def populate(self, pm):
            pm.pmb = self

    class _PM:  # stands in for create_module_pass_manager (reference simulator.py:254-256)
        pmb = None

        def run(self, mod):
            tm = llvm.Target.from_default_triple().create_target_machine(opt=3)
            lvl = self.pmb.opt_level if self.pmb else 0
            pto = llvm.create_pipeline_tuning_options(speed_level=lvl)
            pb = llvm.create_pass_builder(tm, pto)
            pb.getModulePassManager().run(mod, pb)

    if not hasattr(llvm, "create_pass_manager_builder"):
        llvm.create_pass_manager_builder = _PMB
    if not hasattr(llvm, "create_module_pass_manager"):
        llvm.create_module_pass_manager = _PM
    llvm._mcb200_shimmed = True


_cached = None


def load():
    """Return (ref_descriptors_module, ref_simulator_module) of the reference itself.

    The reference package is called ``core``; our drop-in package is importable as
    ``mcircuit_b200.core`` so the two never collide unless the caller put
    ``mcircuit_b200/`` on sys.path - in that case we import the reference under a
    temporary sys.modules swap.
    """
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference not available at %s" % REFERENCE_ROOT)
    _patch_llvmlite()
    saved = {k: v for k, v in sys.modules.items() if k == "core" or k.startswith("core.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        desc = importlib.import_module("core.descriptors")
        sim = importlib.import_module("core.simulator")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        ref_mods = {k: v for k, v in sys.modules.items() if k == "core" or k.startswith("core.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
    _cached = (desc, sim)
    return _cached


@contextlib.contextmanager
def scratch_cwd():
    """The reference JIT writes ./out.txt on every build (simulator.py:266-267)."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory(prefix="mcref_") as d:
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(old)


def build_jit(root, burst_size=500, map_pins=True):
    _, sim = load()
    with scratch_cwd():
        return sim.JIT(root, burst_size, map_pins)
