"""not-gpu: property test - for ANY generated circuit and ANY engine configuration the compiled
program (numpy test double) equals the oracle on every probe, every lane, every step."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import mcircuit_b200.core.descriptors as MD
from _cpu_runtime import CpuRuntime
from mcircuit_b200 import circuits, compiler
from mcircuit_b200.engine import B200Engine
from oracle import restate

ALL_PRIMS = ("Gate", "Not", "Constant", "Clock", "Counter", "Adder", "Register")


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 10**6), n_nodes=st.integers(2, 28), max_width=st.sampled_from([1, 2, 5, 8, 33, 64]),
       feedback=st.booleans(), hierarchy=st.booleans(), map_pins=st.booleans(),
       keep=st.sampled_from(["all", "ports"]), mode=st.sampled_from(["resident", "levels", "levels_grouped", "persistent"]),
       back_edges=st.sampled_from(["auto", "order", "latch"]), share=st.booleans(), sort=st.booleans(),
       tile=st.sampled_from([None, 16]), batch=st.sampled_from([4, 8]))
def test_compiled_program_equals_oracle(seed, n_nodes, max_width, feedback, hierarchy, map_pins, keep, mode,
                                        back_edges, share, sort, tile, batch):
    kw = dict(seed=seed, n_nodes=n_nodes, max_width=max_width, feedback=feedback, hierarchy=hierarchy,
              primitives=ALL_PRIMS)
    root, probes = circuits.random_circuit(MD, **kw)
    lanes = 64 * 3
    ora = restate.RestatedJIT(root, 2, map_pins, lanes=lanes)
    eng = B200Engine(root, 2, map_pins, lanes=lanes, runtime=CpuRuntime(), keep=keep, mode=mode,
                     back_edges=back_edges, share_edge_detectors=share, sort_level_by_fanin=sort,
                     tile_words=tile, resident_batch=batch)
    rng = np.random.default_rng(seed)
    d = eng.program.design
    for p in ("/in0/pin", "/in1/pin"):
        w = d.pin_width[d.find_pin(p)]
        v = rng.integers(0, (1 << w) - 1, size=lanes, dtype=np.uint64, endpoint=True)
        ora.set_pin_state(p, v)
        eng.set_pin_state(p, v)
    for step in range(5):
        ora.step()
        eng.step()
        for p in probes:
            try:
                got = eng.get_pin_state(p, lanes=True)
            except compiler.PinNotObservable:
                assert keep == "ports"
                continue
            assert np.array_equal(got, ora.get_pin_state(p, lanes=True)), (step, p)
    ora.burst()
    eng.burst()
    for p in probes:
        try:
            assert np.array_equal(eng.get_pin_state(p, lanes=True), ora.get_pin_state(p, lanes=True)), p
        except compiler.PinNotObservable:
            pass


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 10**6), n_nodes=st.integers(2, 24), max_width=st.sampled_from([1, 3, 8, 31, 64]),
       map_pins=st.booleans())
def test_c_oracle_equals_numpy_oracle(seed, n_nodes, max_width, map_pins):
    """The two restatements of the reference's algorithm (value-level numpy, bit-sliced C) agree."""
    from oracle import cref
    root, probes = circuits.random_circuit(MD, seed, n_nodes=n_nodes, max_width=max_width, primitives=ALL_PRIMS)
    lanes = 128
    a = restate.RestatedJIT(root, 2, map_pins, lanes=lanes)
    b = cref.CRef(root, 2, map_pins, lanes=lanes, threads=2)
    rng = np.random.default_rng(seed)
    for p in ("/in0/pin", "/in1/pin"):
        v = rng.integers(0, (1 << a._width[p]) - 1, size=lanes, dtype=np.uint64, endpoint=True)
        a.set_pin_state(p, v)
        b.set_pin_state(p, v)
    for step in range(4):
        a.step()
        b.step()
        for p in probes:
            assert np.array_equal(a.get_pin_state(p, lanes=True), b.get_pin_state(p, lanes=True)), (step, p)


@settings(max_examples=36, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 10**6), n_nodes=st.integers(2, 40), max_width=st.sampled_from([1, 2, 5, 8, 33, 64]),
       feedback=st.booleans(), hierarchy=st.booleans(), map_pins=st.booleans(), keep=st.sampled_from(["all", "ports"]),
       share=st.booleans(), guard=st.booleans(), patterns=st.booleans(), stash=st.booleans(), ring=st.booleans(),
       prefix=st.booleans(), model=st.sampled_from(["early", "late"]), cta_cols=st.sampled_from([1, 2, 128]),
       chunks=st.lists(st.integers(1, 40), min_size=1, max_size=4))
def test_serial_stream_equals_oracle(seed, n_nodes, max_width, feedback, hierarchy, map_pins, keep, share, guard, patterns,
                                     stash, ring, prefix, model, cta_cols, chunks):
    """The stream program (K3) under the test double's producer / consumer model - staged rows fetched as early as the
    kernel may with only fenced stores visible, or as late as possible - for ANY circuit, ANY combination of the stream's
    optimisations and ANY split of the run into launches."""
    root, probes = circuits.random_circuit(MD, seed=seed, n_nodes=n_nodes, max_width=max_width, feedback=feedback,
                                           hierarchy=hierarchy, primitives=ALL_PRIMS + ("Register", "Counter"))
    lanes = 64 * 3
    rt = CpuRuntime()
    rt.serial_fetch_model, rt.serial_cta_cols = model, cta_cols
    ora = restate.RestatedJIT(root, 2, map_pins, lanes=lanes)
    eng = B200Engine(root, 2, map_pins, lanes=lanes, runtime=rt, keep=keep, mode="serial", share_edge_detectors=share,
                     serial_guard=guard, adder_prefix=prefix, serial_options=dict(patterns=patterns, stash=stash, ring=ring))
    rng = np.random.default_rng(seed)
    d = eng.program.design
    for p in ("/in0/pin", "/in1/pin"):
        w = d.pin_width[d.find_pin(p)]
        v = rng.integers(0, (1 << w) - 1, size=lanes, dtype=np.uint64, endpoint=True)
        ora.set_pin_state(p, v)
        eng.set_pin_state(p, v)
    for n in chunks:
        eng.run(n)                                        # ONE launch of n steps: the producer runs ahead across steps
        for _ in range(n):
            ora.step()
        for p in probes:
            try:
                got = eng.get_pin_state(p, lanes=True)
            except compiler.PinNotObservable:
                assert keep == "ports"
                continue
            assert np.array_equal(got, ora.get_pin_state(p, lanes=True)), (n, p)
