"""not-gpu: the oracle (oracle/restate.py) against the committed golden vectors that
oracle/gen_golden.py harvested from the reference's own LLVM JIT."""
import numpy as np
import pytest

import mcircuit_b200.core.descriptors as MD
from _helpers import (build, golden, lane_trick_planes, oracle_factory, run_register_case,
                      run_trace_case)
from oracle import restate

TRACES = golden("traces.json")


@pytest.mark.parametrize("case", TRACES, ids=[c["name"] for c in TRACES])
@pytest.mark.parametrize("map_pins", [True, False])
def test_oracle_traces(case, map_pins):
    run_trace_case(case, lambda root, bs: oracle_factory(root, bs, lanes=1, map_pins=map_pins))


def test_examples_goldens_match_survey():
    by = {c["name"]: c for c in TRACES}
    assert [r[0] for r in by["examples_clock"]["values"]] == [1, 0, 1, 0, 1, 0, 1, 0, 1, 0]
    assert by["examples_counter"]["values"][:4] == [[1, 1, 1], [0, 1, 0], [1, 0, 1], [0, 0, 0]]
    for seed in range(4):
        v = by["examples_raw_counter_seed%d" % seed]["values"]
        assert [r[0] | (r[1] << 1) for r in v[:10]] == [0, 0, 1, 1, 2, 2, 3, 3, 0, 0]


@pytest.mark.parametrize("case", golden("adder.json"), ids=lambda c: c["name"])
def test_oracle_adder(case):
    w = case["width"]
    top = MD.Composite()
    top.add_child("add", MD.Adder(w))
    vecs = case["vectors"]
    lanes = len(vecs)
    sim = restate.RestatedJIT(top, 1, True, lanes=lanes)
    sim.set_pin_state("/add/a", np.array([v[0] for v in vecs], dtype=np.uint64))
    sim.set_pin_state("/add/b", np.array([v[1] for v in vecs], dtype=np.uint64))
    sim.set_pin_state("/add/cin", np.array([v[2] for v in vecs], dtype=np.uint64))
    sim.step()
    s = sim.get_pin_state("/add/sum", lanes=True)
    c = sim.get_pin_state("/add/cout", lanes=True)
    for i, (ws, wc) in enumerate(case["results"]):
        assert (int(s[i]), int(c[i])) == (ws, wc), (w, vecs[i])


@pytest.mark.parametrize("case", golden("lane_trick.json"), ids=lambda c: c["name"])
def test_oracle_lane_trick(case):
    """width-1 circuit on 64*n_words lanes == the reference's width-64 JIT words."""
    ins, outs = lane_trick_planes(case)
    lanes = 64 * case["n_words"]
    sim = restate.RestatedJIT(build(case), 1, True, lanes=lanes)
    for p, words in ins.items():
        sim.set_pin_state(p, restate.unpack_lanes(words[None, :]))
    sim.step()
    for p, words in outs.items():
        got = restate.pack_lanes(sim.get_pin_state(p, lanes=True), 1)[0]
        assert np.array_equal(got, words), p


@pytest.mark.parametrize("case", golden("register_intent.json"), ids=lambda c: c["name"])
def test_oracle_register_intent(case):
    run_register_case(case, lambda root, bs: oracle_factory(root, bs, lanes=1))


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    for width in (1, 5, 64):
        hi = (1 << width) - 1
        v = rng.integers(0, hi, size=256, dtype=np.uint64, endpoint=True)
        assert np.array_equal(restate.unpack_lanes(restate.pack_lanes(v, width)), v)


def test_splitmix_known_answers():
    # splitmix64 reference vectors (seed 0 stream: 0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, ...)
    x = np.uint64(0)
    assert int(restate.splitmix64(x)) == 0xE220A8397B1DCDAF
    assert int(restate.splitmix64(np.uint64(0x9E3779B97F4A7C15))) == 0x6E789E6AA1B965F4
