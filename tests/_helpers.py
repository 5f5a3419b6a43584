"""Shared test helpers: golden fixtures and executor-agnostic case runners."""
import json
import os

import numpy as np

import mcircuit_b200.core.descriptors as MD
from mcircuit_b200 import circuits
from oracle import restate

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
U64 = np.uint64


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def build(case, M=MD):
    made = getattr(circuits, case["builder"])(M, **case["args"])
    return made[0] if isinstance(made, tuple) else made


def oracle_factory(root, burst_size, lanes=64, map_pins=True):
    return restate.RestatedJIT(root, burst_size, map_pins, lanes=lanes)


def run_trace_case(case, factory, skip_unobservable=False):
    """factory(root, burst_size) -> executor.  Compares every probe after every step."""
    sim = factory(build(case), case["burst_size"])
    for p, v in case["pokes"]:
        sim.set_pin_state(p, v)
    for step, row in enumerate(case["values"]):
        sim.step()
        for p, want in zip(case["probes"], row):
            try:
                got = sim.get_pin_state(p)
            except KeyError:
                if skip_unobservable:
                    continue
                raise
            assert got == want, (case["name"], "step", step, p, got, want)
    for k, row in enumerate(case["after_bursts"]):
        sim.burst()
        for p, want in zip(case["probes"], row):
            assert sim.get_pin_state(p) == want, (case["name"], "burst", k, p)


def register_shift(M=MD, width=1, n=4):
    top = M.Composite()
    top.add_child("clk", M.Clock())
    top.add_child("d", M.ExposedPin(M.ExposedPin.IN, width))
    for i in range(n):
        top.add_child(f"r{i}", M.Register(width))
    for i in range(n):
        top.connect("clk", "out", f"r{i}", "clock")
        top.connect("d" if i == 0 else f"r{i - 1}", "" if i == 0 else "out", f"r{i}", "data")
    return top


def run_register_case(case, factory):
    if case["kind"] == "register_intent":
        sim = factory(register_shift(MD, case["width"], case["n"]), 3)
        for step, (d, row) in enumerate(zip(case["d_values"], case["values"])):
            sim.set_pin_state("/d/pin", d)
            sim.step()
            for p, want in zip(case["probes"], row):
                assert sim.get_pin_state(p) == want, (case["name"], step, p)
    else:
        root, probes = circuits.random_circuit(
            MD, case["seed"], n_nodes=case["n_nodes"],
            primitives=("Gate", "Not", "Constant", "Clock", "Counter", "Adder", "Register", "Register"))
        assert probes == case["probes"]
        sim = factory(root, 3)
        for step, row in enumerate(case["values"]):
            sim.step()
            for p, want in zip(probes, row):
                assert sim.get_pin_state(p) == want, (case["name"], step, p)


def lane_trick_planes(case):
    """inputs/outputs of a lane-trick case as uint64 arrays [n_words] per 1-bit pin."""
    ins = {p: np.array(v, dtype=U64) for p, v in case["inputs"].items()}
    outs = {p: np.array(v, dtype=U64) for p, v in case["outputs"].items()}
    return ins, outs


def compare_engine_to_oracle(root_builder, factory_engine, lanes, steps, probes, pokes=(),
                             lane_inputs=None, seed=0, burst_size=3):
    """Run engine and oracle on the same per-lane stimulus; compare all lanes of all probes
    after every step.  lane_inputs: list of (pin, width) driven with random per-lane values."""
    rng = np.random.default_rng(seed)
    root = root_builder(MD)
    if isinstance(root, tuple):
        root = root[0]
    ora = restate.RestatedJIT(root, burst_size, True, lanes=lanes)
    eng = factory_engine(root, burst_size, lanes)
    for p, v in pokes:
        ora.set_pin_state(p, v)
        eng.set_pin_state(p, v)
    for pin, width in (lane_inputs or ()):
        hi = (1 << width) - 1
        vals = rng.integers(0, hi, size=lanes, dtype=U64, endpoint=True)
        ora.set_pin_state(pin, vals)
        eng.set_pin_state(pin, vals)
    for step in range(steps):
        ora.step()
        eng.step()
        for p in probes:
            want = ora.get_pin_state(p, lanes=True)
            got = eng.get_pin_state(p, lanes=True)
            assert np.array_equal(got, want), (step, p, got[:4], want[:4])
    return eng, ora


def run_schematic_case(case, factory, root=None):
    """GUI-shaped sheet (tests/golden/schematic.json, values from the reference JIT): pokes, step, probes."""
    if root is None:
        root = circuits.gui_circuit(MD, case["sheet"])
    sim = factory(root, case["burst_size"])
    for step, (pokes, row) in enumerate(zip(case["script"], case["values"])):
        for p, v in pokes:
            sim.set_pin_state(p, v)
        sim.step()
        for p, want in zip(case["probes"], row):
            assert sim.get_pin_state(p) == want, (case["name"], step, p)
    if "after_burst" in case:
        sim.burst()
        for p, want in zip(case["probes"], case["after_burst"]):
            assert sim.get_pin_state(p) == want, (case["name"], "burst", p)
    return sim
