"""not-gpu: the plain-C restatement (oracle/mcref.c) against the golden vectors from the
reference JIT and against the numpy restatement."""
import numpy as np
import pytest

import mcircuit_b200.core.descriptors as MD
from _helpers import build, golden, lane_trick_planes, run_register_case, run_trace_case
from mcircuit_b200 import circuits
from oracle import cref, restate

TRACES = golden("traces.json")


@pytest.mark.parametrize("case", TRACES, ids=[c["name"] for c in TRACES])
@pytest.mark.parametrize("map_pins", [True, False])
def test_cref_traces(case, map_pins):
    run_trace_case(case, lambda root, bs: cref.CRef(root, bs, map_pins, lanes=64, threads=1))


@pytest.mark.parametrize("case", golden("register_intent.json"), ids=lambda c: c["name"])
def test_cref_register_intent(case):
    run_register_case(case, lambda root, bs: cref.CRef(root, bs, True, lanes=64, threads=1))


@pytest.mark.parametrize("case", golden("adder.json"), ids=lambda c: c["name"])
def test_cref_adder(case):
    w = case["width"]
    top = MD.Composite()
    top.add_child("add", MD.Adder(w))
    vecs = case["vectors"]
    lanes = (len(vecs) + 63) // 64 * 64
    pad = lanes - len(vecs)
    sim = cref.CRef(top, 1, True, lanes=lanes, threads=2)
    col = lambda k: np.array([v[k] for v in vecs] + [0] * pad, dtype=np.uint64)
    sim.set_pin_state("/add/a", col(0))
    sim.set_pin_state("/add/b", col(1))
    sim.set_pin_state("/add/cin", col(2))
    sim.step()
    s, c = sim.get_pin_state("/add/sum", lanes=True), sim.get_pin_state("/add/cout", lanes=True)
    for i, (ws, wc) in enumerate(case["results"]):
        assert (int(s[i]), int(c[i])) == (ws, wc), (w, vecs[i])


@pytest.mark.parametrize("case", golden("lane_trick.json"), ids=lambda c: c["name"])
def test_cref_lane_trick(case):
    ins, outs = lane_trick_planes(case)
    sim = cref.CRef(build(case), 1, True, lanes=64 * case["n_words"], threads=3)
    for p, words in ins.items():
        sim.set_planes(p, words[None, :])
    sim.step()
    for p, words in outs.items():
        assert np.array_equal(sim.planes(p)[0], words), p


@pytest.mark.parametrize("seed", range(300, 320))
def test_cref_equals_numpy_restatement_on_all_lanes(seed):
    root, probes = circuits.random_circuit(
        MD, seed, n_nodes=20,
        primitives=("Gate", "Not", "Constant", "Clock", "Counter", "Adder", "Register"))
    a = restate.RestatedJIT(root, 2, True, lanes=192)
    b = cref.CRef(root, 2, True, lanes=192, threads=2)
    rng = np.random.default_rng(seed)
    for p in ("/in0/pin", "/in1/pin"):
        v = rng.integers(0, (1 << a._width[p]) - 1, size=192, dtype=np.uint64, endpoint=True)
        a.set_pin_state(p, v)
        b.set_pin_state(p, v)
    for step in range(5):
        a.step()
        b.step()
        for p in probes:
            assert np.array_equal(a.get_pin_state(p, lanes=True), b.get_pin_state(p, lanes=True)), (step, p)
    a.burst()
    b.burst()
    for p in probes:
        assert np.array_equal(a.get_pin_state(p, lanes=True), b.get_pin_state(p, lanes=True)), p
