"""Headless stand-in for the reference GUI's ``Schematic.reconstruct`` (``diagram.py:88-119``).

The GUI is out of scope, but it is the *caller* of the simulation path: Start/Stop does
``JIT(schematic.reconstruct(), 500, True)`` (``app.py:722-735``).  ``reconstruct`` clears the
schematic's composite, adds every element as a child in element order, collects the output
pins (``sources``) and input pins (``dests``) that touch a wire, and calls ``connect`` for
every (source, dest) pair - sources outer, dests inner (``itertools.product``) - whose grid
points are joined by a wire path.  Geometry (PySide6) only decides *which* pins share a
wire; here that is given directly as nets, and the composite that comes out - children
order, edge insertion order, connections - is the one ``reconstruct`` would build, which is
all the engine's evaluation order depends on.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

from .core import descriptors as _own

PinRef = Tuple[str, str]        # (element name, pin name)


def element_inputs(desc) -> List[str]:
    """Input pin names in ``Element.all_inputs`` order (diagram.py:46-64)."""
    t = type(desc).__name__
    if t == "Not":
        return ["in"]
    if t == "Gate":
        return ["in%d" % i for i in range(desc.num_inputs)]
    if t == "ExposedPin":
        return ["pin"] if desc.direction == 1 else []
    if hasattr(desc, "graph"):
        return [name for name, _ in desc.all_inputs()]
    return [name for name, _ in desc.all_inputs()]       # primitives the GUI library lacks


def element_outputs(desc) -> List[str]:
    """Output pin names in ``Element.all_outputs`` order (diagram.py:66-78)."""
    t = type(desc).__name__
    if t in ("Not", "Gate"):
        return ["out"]
    if t == "ExposedPin":
        return ["pin"] if desc.direction == 0 else []
    return [name for name, _ in desc.all_outputs()]


def reconstruct(elements: Sequence[Tuple[str, object]], nets: Iterable[Iterable[PinRef]], M=_own):
    """elements: (name, descriptor) in schematic order; nets: groups of pins joined by wires.
    Returns the Composite ``Schematic.reconstruct`` would produce."""
    s = M.Composite()
    net_of = {}
    for k, net in enumerate(nets):
        for ref in net:
            net_of[tuple(ref)] = k
    sources, dests = [], []
    for name, desc in elements:
        s.add_child(name, desc)
        for pin in element_outputs(desc):
            if (name, pin) in net_of:                     # "if not self.wires.has_node(p): return"
                sources.append((name, pin))
        for pin in element_inputs(desc):
            if (name, pin) in net_of:
                dests.append((name, pin))
    for src in sources:                                   # itertools.product(sources, dests)
        for dst in dests:
            if net_of[src] == net_of[dst]:                # nx.has_path(self.wires, ...)
                s.connect(src[0], src[1], dst[0], dst[1])
    return s
