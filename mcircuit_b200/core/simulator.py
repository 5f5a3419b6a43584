"""Drop-in for mcircuit's ``core.simulator`` (``/root/reference/core/simulator.py``).

``from core.simulator import JIT`` keeps working when ``mcircuit_b200/`` is on ``sys.path``
(or ``from mcircuit_b200.core.simulator import JIT``): ``JIT`` is the B200 engine.  The
module-level helpers the reference exports - and that its GUI imports (``app.py:5``) - are
re-provided with identical outputs but linear cost (the reference's are
O(nodes x connections)).
"""
try:
    from ..compiler import _SCHED_EMIT, flatten
    from ..engine import B200Engine, Executor
except ImportError:
    # drop-in mode: ``mcircuit_b200/`` itself is on sys.path, so this module is the top-level
    # ``core.simulator`` exactly like the reference's; reach the rest of the package by name.
    # (The compiler is duck-typed on descriptor class names, so descriptors imported as
    # ``core.descriptors`` and as ``mcircuit_b200.core.descriptors`` both work.)
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
    from mcircuit_b200.compiler import _SCHED_EMIT, flatten
    from mcircuit_b200.engine import B200Engine, Executor

__all__ = ["Executor", "B200Engine", "JIT", "iter_simulation_pins",
           "iter_simulation_connections", "iter_simulation_topology", "_map_to_sources"]


def iter_simulation_pins(node, path="/"):
    """(abs_pin_path, width, 'in'|'out'|'internal') for every leaf pin
    (reference simulator.py:16-25)."""
    d = flatten(node, True)
    for p, w, k in d.iter_pins():
        yield path + p[1:], w, k


def iter_simulation_connections(node, path="/"):
    """(abs_src_pin, abs_dst_pin) for every connection (reference simulator.py:28-38)."""
    d = flatten(node, True)
    for s, t in zip(d.conn_src, d.conn_dst):
        yield path + d.pin_path(s)[1:], path + d.pin_path(t)[1:]


def iter_simulation_topology(node, path="/"):
    """The evaluation schedule: ('emit', (descriptor, leaf_path)) and
    ('propagate', (src_pin, dst_pin)) in the reference's order (simulator.py:41-53)."""
    d = flatten(node, True)
    for kind, arg in d.schedule:
        if kind == _SCHED_EMIT:
            yield "emit", (d.leaf_desc[arg], path + d.leaf_path[arg][1:])
        else:
            yield "propagate", (path + d.pin_path(d.conn_src[arg])[1:],
                                path + d.pin_path(d.conn_dst[arg])[1:])


def _map_to_sources(desc):
    """pin -> ultimate source pin, or None when undriven (reference simulator.py:56-74)."""
    d = flatten(desc, True)
    out = {}
    for pin in range(len(d.pin_width)):
        s = int(d.pin_src[pin])
        out[d.pin_path(pin)] = None if s < 0 else d.pin_path(s)
    return out


class JIT(B200Engine):
    """Same positional constructor as the reference: ``JIT(root, burst_size, map_pins)``."""
