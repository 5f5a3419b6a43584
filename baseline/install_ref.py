"""Put the UNMODIFIED reference where the benchmark can run it on the GPU box.

The reference (matijakevic/mcircuit) is pure Python with no packaging metadata (no setup.py /
pyproject.toml; requirements.txt is UTF-16), so `pip install --target baseline/_ref /root/reference`
has nothing to install.  What the base contract's install step is for - the stock reference, runnable
from `baseline/_ref`, travelling with the gpurun snapshot - is done here by copying the files of the
simulation path verbatim: core/ (descriptors + the llvmlite JIT), examples/ (the benchmark scripts
BASELINE.json's configs[0] names) and LICENSE.  `baseline/_ref/` is git-ignored (never part of this
repo's history) and not gpurun-ignored.  Called from `__graft_entry__.build()` whenever
/root/reference exists (this container); on the GPU box the copy that travelled is used as is."""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("MCIRCUIT_REFERENCE_SOURCE", "/root/reference")
WANTED = ("core", "examples")


def installed() -> bool:
    return os.path.isfile(os.path.join(DEST, "core", "simulator.py"))


def install(force=False) -> str:
    """Returns 'installed', 'present' or 'unavailable: <why>'."""
    if installed() and not force:
        return "present"
    if not os.path.isfile(os.path.join(SOURCE, "core", "simulator.py")):
        return "unavailable: %s has no core/simulator.py" % SOURCE
    os.makedirs(DEST, exist_ok=True)
    for sub in WANTED:
        src, dst = os.path.join(SOURCE, sub), os.path.join(DEST, sub)
        if os.path.isdir(src):
            if os.path.isdir(dst):
                shutil.rmtree(dst)
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for name in ("LICENSE", "README.md"):
        if os.path.isfile(os.path.join(SOURCE, name)):
            shutil.copy2(os.path.join(SOURCE, name), os.path.join(DEST, name))
    return "installed"


if __name__ == "__main__":
    print(install(force=True))
