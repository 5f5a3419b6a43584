"""not-gpu: unusual but legal wiring, three ways - the live reference JIT (when present), the
oracle, and the compiled program on the numpy test double."""
import pytest

import mcircuit_b200.core.descriptors as MD
from _cpu_runtime import CpuRuntime
from mcircuit_b200.engine import B200Engine
from oracle import restate


def input_pin_as_source(M):
    """connect(a.in0 -> b.in0): b.in0 aliases a's INPUT pin (an undriven net both read)."""
    top = M.Composite()
    top.add_child("x", M.ExposedPin(M.ExposedPin.IN, 4))
    top.add_child("a", M.Gate(M.Gate.AND, 4))
    top.add_child("b", M.Gate(M.Gate.XOR, 4))
    top.connect("a", "in0", "b", "in0")
    top.connect("x", "", "a", "in1")
    top.connect("a", "out", "b", "in1")
    return top, ["/a/out", "/b/out", "/b/in0", "/a/in0"], [("/a/in0", 0b1100), ("/x/pin", 0b1010)]


def output_pin_as_destination(M):
    """connect(x.pin -> g.out): g's output pin aliases the port's net, so g WRITES the port."""
    top = M.Composite()
    top.add_child("x", M.ExposedPin(M.ExposedPin.IN, 3))
    top.add_child("g", M.Not(3))
    top.add_child("h", M.Not(3))
    top.connect("x", "", "g", "out")
    top.connect("x", "", "h", "in")
    return top, ["/x/pin", "/g/out", "/h/out"], [("/g/in", 0b101)]


def constant_through_ports(M):
    inner = M.Composite()
    inner.add_child("i", M.ExposedPin(M.ExposedPin.IN, 5))
    inner.add_child("n", M.Not(5))
    inner.add_child("o", M.ExposedPin(M.ExposedPin.OUT, 5))
    inner.add_child("dangling", M.ExposedPin(M.ExposedPin.OUT, 2))
    inner.connect("i", "", "n", "in")
    inner.connect("n", "out", "o", "")
    mid = M.Composite()
    mid.add_child("i", M.ExposedPin(M.ExposedPin.IN, 5))
    mid.add_child("u", inner)
    mid.add_child("o", M.ExposedPin(M.ExposedPin.OUT, 5))
    mid.connect("i", "", "u", "i")
    mid.connect("u", "o", "o", "")
    top = M.Composite()
    top.add_child("c", M.Constant(5, 0b10110))
    top.add_child("m", mid)
    top.add_child("m2", mid)
    top.add_child("z", M.ExposedPin(M.ExposedPin.OUT, 5))
    top.connect("c", "out", "m", "i")
    top.connect("m", "o", "m2", "i")
    top.connect("m2", "o", "z", "")
    return top, ["/z/pin", "/m/u/n/out", "/m2/u/o/pin", "/m/u/dangling/pin", "/c/out"], []


def feedback_through_ports(M):
    """A ring of three inverters, each inside its own composite: one back edge through ports."""
    inv = M.Composite()
    inv.add_child("a", M.ExposedPin(M.ExposedPin.IN))
    inv.add_child("n", M.Not())
    inv.add_child("y", M.ExposedPin(M.ExposedPin.OUT))
    inv.connect("a", "", "n", "in")
    inv.connect("n", "out", "y", "")
    top = M.Composite()
    for k in range(3):
        top.add_child(f"r{k}", inv)
    top.add_child("tap", M.ExposedPin(M.ExposedPin.OUT))
    for k in range(3):
        top.connect(f"r{k}", "y", f"r{(k + 1) % 3}", "a")
    top.connect("r0", "y", "tap", "")
    return top, ["/tap/pin", "/r1/n/out", "/r2/y/pin"], []


def counter_clocked_by_logic(M):
    """Counters on a derived clock; one of them is emitted before the gate that makes it."""
    top = M.Composite()
    top.add_child("late", M.Counter(3))
    top.add_child("clk", M.Clock())
    top.add_child("en", M.ExposedPin(M.ExposedPin.IN))
    top.add_child("g", M.Gate(M.Gate.AND))
    top.add_child("c1", M.Counter(2))
    top.add_child("c2", M.Counter(4))
    top.connect("clk", "out", "g", "in0")
    top.connect("en", "", "g", "in1")
    top.connect("g", "out", "c1", "clock")
    top.connect("g", "out", "c2", "clock")
    top.connect("c1", "out", "late", "clock") if False else top.connect("g", "out", "late", "clock")
    return top, ["/g/out", "/c1/out", "/c2/out", "/late/out", "/late/prevclock", "/c1/prevclock"], [("/en/pin", 1)]


def two_primitives_store_to_one_net(M):
    """b.out is the destination of a.out: both primitives store to one global, one after the
    other, every step (the reference's globals simply take the writes in emit order)."""
    top = M.Composite()
    top.add_child("x", M.ExposedPin(M.ExposedPin.IN, 4))
    top.add_child("a", M.Not(4))
    top.add_child("b", M.Gate(M.Gate.XOR, 4))
    top.add_child("r", M.Not(4))
    top.add_child("late", M.Gate(M.Gate.OR, 4))
    top.connect("x", "", "a", "in")
    top.connect("a", "out", "b", "out")
    top.connect("x", "", "b", "in0")
    top.connect("a", "out", "r", "in")
    top.connect("b", "out", "late", "in0")
    top.connect("r", "out", "late", "in1")
    return top, ["/a/out", "/b/out", "/r/out", "/late/out"], [("/x/pin", 0b0110), ("/b/in1", 0b1001)]


CASES = [two_primitives_store_to_one_net, input_pin_as_source, output_pin_as_destination, constant_through_ports, feedback_through_ports,
         counter_clocked_by_logic]


def _engines(build, map_pins):
    root, probes, pokes = build(MD)
    sims = [restate.RestatedJIT(root, 3, map_pins)]
    for mode in ("resident", "levels"):
        for kw in ({}, {"back_edges": "latch"}, {"share_edge_detectors": True}, {"keep": "all", "sort_level_by_fanin": True}):
            sims.append(B200Engine(root, 3, map_pins, runtime=CpuRuntime(), mode=mode, **kw))
    return sims, probes, pokes


@pytest.mark.parametrize("build", CASES, ids=lambda f: f.__name__)
@pytest.mark.parametrize("map_pins", [True, False])
def test_engine_equals_oracle(build, map_pins):
    sims, probes, pokes = _engines(build, map_pins)
    for step in range(9):
        if step == 4:
            for s in sims:
                for p, v in pokes:
                    s.set_pin_state(p, v)
        for s in sims:
            s.step()
        want = [sims[0].get_pin_state(p) for p in probes]
        for s in sims[1:]:
            assert [s.get_pin_state(p) for p in probes] == want, (build.__name__, step)


@pytest.mark.parametrize("build", CASES, ids=lambda f: f.__name__)
@pytest.mark.parametrize("map_pins", [True, False])
def test_oracle_equals_live_reference(reference, build, map_pins):
    from oracle import refshim
    D, _ = reference
    root, probes, pokes = build(D)
    jit = refshim.build_jit(root, 3, map_pins)
    ora = restate.RestatedJIT(build(MD)[0], 3, map_pins)
    for step in range(9):
        if step == 4:
            for p, v in pokes:
                jit.set_pin_state(p, v)
                ora.set_pin_state(p, v)
        jit.step()
        ora.step()
        assert [jit.get_pin_state(p) for p in probes] == [ora.get_pin_state(p) for p in probes], (build.__name__, step)
    jit.burst()
    ora.burst()
    assert [jit.get_pin_state(p) for p in probes] == [ora.get_pin_state(p) for p in probes]
