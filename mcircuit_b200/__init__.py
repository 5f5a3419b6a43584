"""mcircuit_b200 - B200-native gate-level evaluation engine, drop-in for the simulation
path of matijakevic/mcircuit (descriptors + Executor).  See DESIGN.md."""
from .core import descriptors  # noqa: F401
from .core.descriptors import (Adder, Clock, Composite, Constant, Counter, ExposedPin,  # noqa: F401
                               Gate, Not, Register)
from .compiler import CircuitError, PinNotObservable, compile_circuit  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name in ("B200Engine", "Executor", "JIT"):
        from .core import simulator
        return getattr(simulator, name)
    raise AttributeError(name)
