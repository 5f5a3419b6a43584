"""Circuit descriptors - the drop-in surface for mcircuit's ``core.descriptors``.

Same class names, constructor signatures, attribute names and generator outputs as the
reference (``/root/reference/core/descriptors.py:6-180``) so that circuits written against
mcircuit build unchanged.  (The reference's JIT dispatches on its own classes, so results are
pinned by building the same circuit twice, once with each module - see
``mcircuit_b200.circuits``.)  Nothing here touches a GPU.

Pin tables are data (``_INPUTS/_OUTPUTS/_INTERNALS`` specs resolved against ``self``)
rather than one hand-written generator per class; the observable behaviour is the
reference's: ``all_inputs/all_outputs/all_internals`` yield ``(pin_name, width)`` in the
same order (descriptors.py:10-22).
"""
from copy import deepcopy

import networkx as nx

__all__ = [
    "Descriptor", "ExposedPin", "Constant", "Not", "Gate", "Register", "Counter",
    "Clock", "Adder", "Composite",
]


def _resolve(spec, obj):
    # spec entries: (pin_name, width) where width is an int or the name of an attribute
    for pin, width in spec:
        yield pin, (getattr(obj, width) if isinstance(width, str) else width)


class Descriptor:
    """Base of every descriptor (reference descriptors.py:6-22)."""

    _INPUTS = ()
    _OUTPUTS = ()
    _INTERNALS = ()

    def _set(self, **attributes):
        """Constructor helper: descriptors are plain attribute bags (the GUI's property editors,
        editors.py, read and write these attributes by name)."""
        vars(self).update(attributes)

    def clone(self):
        return deepcopy(self)

    def all_inputs(self):
        return _resolve(self._INPUTS, self)

    def all_outputs(self):
        return _resolve(self._OUTPUTS, self)

    def all_internals(self):
        return _resolve(self._INTERNALS, self)

    def all_pins(self):
        for group in (self.all_inputs(), self.all_outputs(), self.all_internals()):
            yield from group


class ExposedPin(Descriptor):
    """A port (reference descriptors.py:25-39).  An IN port *drives* the inside of its
    composite, so it exposes an output pin called ``pin``; an OUT port exposes an input
    pin called ``pin``."""

    IN = 0
    OUT = 1

    def __init__(self, direction, width=1):
        self._set(direction=direction, width=width)

    def all_inputs(self):
        return iter((("pin", self.width),) if self.direction == ExposedPin.OUT else ())

    def all_outputs(self):
        return iter((("pin", self.width),) if self.direction == ExposedPin.IN else ())


class Constant(Descriptor):
    """``out`` holds ``value``; written once after build (reference descriptors.py:42-49,
    simulator.py:220-221, 275-276)."""

    _OUTPUTS = (("out", "width"),)

    def __init__(self, width=1, value=0):
        self._set(width=width, value=value)


class Not(Descriptor):
    """``out = ~in`` (reference descriptors.py:52-61, simulator.py:77-80)."""

    _INPUTS = (("in", "width"),)
    _OUTPUTS = (("out", "width"),)

    def __init__(self, width=1):
        self._set(width=width)


class Gate(Descriptor):
    """n-input bitwise AND/OR/XOR with optional output negation
    (reference descriptors.py:64-78, simulator.py:83-97)."""

    AND = 0
    OR = 1
    XOR = 2

    _OUTPUTS = (("out", "width"),)

    def __init__(self, op, width=1, num_inputs=2, negated=False):
        self._set(op=op, width=width, num_inputs=num_inputs, negated=negated)

    def all_inputs(self):
        return ((f"in{i}", self.width) for i in range(self.num_inputs))


class Register(Descriptor):
    """Rising-edge register (reference descriptors.py:81-94).  The reference's emitter for
    it does not compile (simulator.py:115-127); the engine implements the
    evident intent and says so."""

    _INPUTS = (("clock", 1), ("data", "width"))
    _OUTPUTS = (("out", "width"),)
    _INTERNALS = (("prevclock", 1),)

    def __init__(self, width=1):
        self._set(width=width)


class Counter(Descriptor):
    """Rising-edge "counter" (reference descriptors.py:97-109).  As emitted by the
    reference only bit 0 ever changes (simulator.py:100-112)."""

    _INPUTS = (("clock", 1),)
    _OUTPUTS = (("out", "width"),)
    _INTERNALS = (("prevclock", 1),)

    def __init__(self, width=1):
        self._set(width=width)


class Clock(Descriptor):
    """Toggles every step (reference descriptors.py:112-122, simulator.py:130-132);
    ``short``/``long``/``count`` exist but are unused by the reference."""

    _OUTPUTS = (("out", 1),)
    _INTERNALS = (("count", 64),)

    def __init__(self):
        self._set(short=1, long=1)


class Adder(Descriptor):
    """``sum = a + b + cin``; ``cout`` is the OR of the two *signed* overflow flags
    (reference descriptors.py:125-137, simulator.py:135-147)."""

    _INPUTS = (("cin", 1), ("a", "width"), ("b", "width"))
    _OUTPUTS = (("cout", 1), ("sum", "width"))

    def __init__(self, width=1):
        self._set(width=width)


class Composite(Descriptor):
    """Hierarchy node (reference descriptors.py:140-180).

    ``graph`` is an ``nx.DiGraph`` whose nodes are child names (attributes ``descriptor``
    and ``label``) and whose edges point from the *destination* child to the *source*
    child; the engine's evaluation order is the DFS post-order of this graph, so node and
    edge insertion order are part of the contract.  ``connections`` is a set
    of ``(src_child, src_pin, dst_child, dst_pin)`` with ports already rewritten by
    ``_translate_pin``.
    """

    def __init__(self):
        self._set(graph=nx.DiGraph(), connections=set())

    # -- pin-name rewriting (reference descriptors.py:146-152) -------------------------
    def _translate_pin(self, child, pin):
        target = self.get_child(child)
        if isinstance(target, ExposedPin):
            return child, "pin"
        if isinstance(target, Composite):
            return f"{child}/{pin}", "pin"
        return child, pin

    def connect(self, child1, pin1, child2, pin2):
        # edge first, on the untranslated names, destination -> source
        self.graph.add_edge(child2, child1)
        src = self._translate_pin(child1, pin1)
        dst = self._translate_pin(child2, pin2)
        self.connections.add(src + dst)

    def add_child(self, name, descriptor):
        self.graph.add_node(name, descriptor=descriptor, label=name)

    def get_child(self, name) -> Descriptor:
        return self.graph.nodes[name]["descriptor"]

    def _ports(self, direction):
        for name, attrs in self.graph.nodes.items():
            d = attrs["descriptor"]
            if isinstance(d, ExposedPin) and d.direction == direction:
                yield name, d.width

    def all_inputs(self):
        return self._ports(ExposedPin.IN)

    def all_outputs(self):
        return self._ports(ExposedPin.OUT)
