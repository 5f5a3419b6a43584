"""``core`` - same package name and modules as the reference (core/descriptors.py,
core/simulator.py).  ``simulator`` is imported lazily so that building descriptors never
needs the CUDA library."""
from . import descriptors  # noqa: F401
