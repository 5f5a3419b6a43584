"""-m gpu: the streamed serial kernel (k_run_stream, K3 v3) through the C ABI against the goldens
harvested from the reference JIT and against the oracle.  Bit-exact: integer work."""
import numpy as np
import pytest

import mcircuit_b200.core.descriptors as MD
from _helpers import (build, compare_engine_to_oracle, golden, lane_trick_planes, run_register_case,
                      run_trace_case)
from mcircuit_b200 import circuits, compiler
from mcircuit_b200.engine import B200Engine
from oracle import cref, restate

pytestmark = pytest.mark.gpu
TRACES = golden("traces.json")
U64 = np.uint64


def serial_factory(keep="all", map_pins=True, **kw):
    def make(root, burst_size, lanes=64):
        return B200Engine(root, burst_size, map_pins, lanes=lanes, mode="serial", keep=keep, **kw)
    return make


@pytest.mark.parametrize("case", TRACES, ids=[c["name"] for c in TRACES])
@pytest.mark.parametrize("keep,map_pins,share,wb", [("all", True, False, 0), ("ports", True, True, 8),
                                                     ("all", False, False, 4), ("ports", True, False, 1)])
def test_gpu_serial_traces(case, keep, map_pins, share, wb):
    run_trace_case(case, serial_factory(keep, map_pins, share_edge_detectors=share, serial_word_bytes=wb),
                   skip_unobservable=(keep == "ports"))


@pytest.mark.parametrize("case", golden("register_intent.json"), ids=lambda c: c["name"])
def test_gpu_serial_register_intent(case):
    run_register_case(case, serial_factory())


@pytest.mark.parametrize("case", golden("adder.json"), ids=lambda c: c["name"])
def test_gpu_serial_adder(case):
    w = case["width"]
    top = MD.Composite()
    top.add_child("add", MD.Adder(w))
    vecs = case["vectors"]
    lanes = (len(vecs) + 63) // 64 * 64
    pad = lanes - len(vecs)
    eng = serial_factory()(top, 1, lanes)
    col = lambda k: np.array([v[k] for v in vecs] + [0] * pad, dtype=np.uint64)
    eng.set_pin_state("/add/a", col(0))
    eng.set_pin_state("/add/b", col(1))
    eng.set_pin_state("/add/cin", col(2))
    eng.step()
    s = eng.get_pin_state("/add/sum", lanes=True)
    c = eng.get_pin_state("/add/cout", lanes=True)
    for i, (ws, wc) in enumerate(case["results"]):
        assert (int(s[i]), int(c[i])) == (ws, wc), (w, vecs[i])


@pytest.mark.parametrize("case", golden("lane_trick.json"), ids=lambda c: c["name"])
@pytest.mark.parametrize("wb", [8, 4, 2])
def test_gpu_serial_lane_trick(case, wb):
    ins, outs = lane_trick_planes(case)
    nw = case["n_words"]
    eng = B200Engine(build(case), 1, True, lanes=64 * nw * 5, mode="serial", keep="ports", serial_word_bytes=wb)
    for p, words in ins.items():
        eng.set_pin_planes(p, np.tile(words, 5)[None, :])
    eng.step()
    for p, words in outs.items():
        assert np.array_equal(eng.get_pin_planes(p)[0], np.tile(words, 5)), p


@pytest.mark.parametrize("seed", range(100, 160))
def test_gpu_serial_vs_oracle_all_lanes(seed):
    prims = ("Gate", "Not", "Constant", "Clock", "Counter", "Counter", "Adder", "Register", "Register")
    root, probes = circuits.random_circuit(MD, seed, n_nodes=26, primitives=prims)
    d = compiler.flatten(root)
    lane_inputs = [(p, d.pin_width[d.find_pin(p)]) for p in ("/in0/pin", "/in1/pin")]
    compare_engine_to_oracle(lambda M: circuits.random_circuit(M, seed, n_nodes=26, primitives=prims),
                             lambda r, bs, lanes: serial_factory(share_edge_detectors=bool(seed % 2),
                                                                 serial_word_bytes=(0, 8, 4, 2, 1)[seed % 5])(r, bs, lanes),
                             lanes=64 * 37, steps=8, probes=probes, lane_inputs=lane_inputs, seed=seed)


@pytest.mark.parametrize("share", [False, True])
@pytest.mark.parametrize("wb", [8, 4, 2])
def test_gpu_serial_c4_shape_vs_oracle(share, wb):
    args = dict(n_lfsr=8, lfsr_bits=16, stages=40, taps=(15, 13, 12, 10))
    root = circuits.lfsr_pipeline(MD, **args)
    lanes = 64 * 67                        # ragged: not a multiple of a warp's columns
    eng = B200Engine(root, 1, True, lanes=lanes, mode="serial", keep="ports", share_edge_detectors=share,
                     serial_word_bytes=wb, observe=["/l3_q7/out", "/p17/out", "/p17/prevclock"])
    ora = cref.CRef(root, 1, True, lanes=lanes)
    rng = np.random.default_rng(4)
    for n in range(8):
        for b in range(16):
            v = rng.integers(0, 2, size=lanes, dtype=U64)
            eng.set_pin_state(f"/l{n}_q{b}/out", v)
            ora.set_pin_state(f"/l{n}_q{b}/out", v)
    probes = [f"/tap{n}/pin" for n in range(8)] + ["/pipe_out/pin", "/l3_q7/out", "/p17/out", "/p17/prevclock"]
    done = 0
    for chunk in (1, 1, 2, 7, 100, 333):
        eng.rt.launch_count(reset=True)
        eng.run(chunk)
        assert eng.rt.launch_count() == 1           # any number of steps: one launch, no host round trip
        ora.run(chunk)
        done += chunk
        for p in probes:
            assert np.array_equal(eng.get_pin_planes(p), ora.planes(p)), (chunk, p)
    if share:
        assert eng.program.serial_info["regions"] >= 1
        assert eng.regions_skipped() == done // 2   # the clock falls every other step


def test_gpu_serial_guard_per_cta_with_lane_dependent_enable():
    """Same circuit as the CPU case, on the device: a CTA (128 thread columns) whose lanes are all idle
    skips the region, a mixed one and the active ones execute it; every lane matches the oracle."""
    from test_serial_cpu import run_gated_clock_case
    for wb in (8, 4):
        cols_per_word = 8 // wb
        lanes = 64 * (128 // cols_per_word) * 5              # five CTAs
        idle = 64 * (128 // cols_per_word) + 64 * 40         # CTA 0 idle, CTA 1 mixed, the rest active
        run_gated_clock_case(lambda root, n: serial_factory("ports", share_edge_detectors=True,
                                                            serial_word_bytes=wb)(root, 1, n), lanes, idle, 12)


@pytest.mark.parametrize("options", [{}, {"stash": False}, {"ring": False}, {"patterns": False}])
@pytest.mark.parametrize("wb", [8, 1])
def test_gpu_serial_many_steps_in_one_launch(options, wb):
    """The producer runs ahead across the end of a step: run(k) in ONE launch against the oracle, every pin,
    every lane; with each optimisation of the stream switched off in turn."""
    cases = [(circuits.lfsr_pipeline(MD, 5, 8, 40, taps=(7, 5, 1, 0)), True), (circuits.counter_circuit(MD), False),
             (circuits.raw_counter(MD, 3, 2), False), (circuits.gate_level_lfsr(MD), False)]
    for seed in (400, 401, 402, 403):
        cases.append((circuits.random_circuit(MD, seed, n_nodes=30, primitives=(
            "Gate", "Not", "Constant", "Clock", "Counter", "Counter", "Adder", "Register", "Register"))[0], bool(seed % 2)))
    lanes = 64 * 150                                         # ragged over two CTAs at 8 bytes per thread
    for root, share in cases:
        eng = B200Engine(root, 3, True, lanes=lanes, mode="serial", keep="all", share_edge_detectors=share,
                         serial_options=options, serial_word_bytes=wb)
        ora = restate.RestatedJIT(root, 3, True, lanes=lanes)
        pins = [p[0] for p in compiler.flatten(root).iter_pins()]
        rng = np.random.default_rng(11)
        for p in pins:
            if p.endswith("/out") and "_q" in p:
                v = rng.integers(0, 2, size=lanes, dtype=U64)
                eng.set_pin_state(p, v)
                ora.set_pin_state(p, v)
        for chunk in (1, 2, 5, 17, 64, 301):
            eng.run(chunk)
            for _ in range(chunk):
                ora.step()
            for p in pins:
                assert np.array_equal(eng.get_pin_state(p, lanes=True), ora.get_pin_state(p, lanes=True)), (chunk, p, options)


def test_gpu_serial_tiled_lanes():
    """Lane tiles (memory budget) with a serial program: every tile runs all steps on its own."""
    root = circuits.lfsr_pipeline(MD, 4, 8, 20, taps=(7, 5, 1, 0))
    lanes = 64 * 96
    eng = B200Engine(root, 1, True, lanes=lanes, mode="serial", keep="ports", tile_words=32, share_edge_detectors=True)
    assert eng.n_tiles == 3
    ora = cref.CRef(root, 1, True, lanes=lanes)
    rng = np.random.default_rng(9)
    for n in range(4):
        v = rng.integers(0, 2, size=lanes, dtype=U64)
        eng.set_pin_state(f"/l{n}_q0/out", v)
        ora.set_pin_state(f"/l{n}_q0/out", v)
    eng.run(41)
    ora.run(41)
    for p in ["/pipe_out/pin"] + [f"/tap{n}/pin" for n in range(4)]:
        assert np.array_equal(eng.get_pin_planes(p), ora.planes(p)), p


@pytest.mark.parametrize("mode,tune", [("serial", None), ("levels", None), ("levels", ("level_tma", 1)), ("resident", None)])
def test_gpu_kernels_never_write_outside_their_lane_columns(mode, tune):
    """Own bounds check (compute-sanitizer is not available on this pool): with a ragged lane count the rows carry
    padding columns up to the 128-byte stride, and the plane ends in a pad; whatever the kernels do with threads
    beyond the last lane column (K3 parks them on a private dummy word), those words must still be zero after a
    run - and the results must still match the oracle."""
    root = circuits.lfsr_pipeline(MD, 6, 16, 48, taps=(15, 13, 12, 10))
    lanes = 64 * 150                                     # 150 lane words in rows of 160
    eng = B200Engine(root, 1, True, lanes=lanes, mode=mode, keep="ports")
    if tune:
        old = eng.rt.set_tuning(*tune)
    try:
        assert eng.stride == 160 and eng.n_tiles == 1
        ora = cref.CRef(root, 1, True, lanes=lanes)
        rng = np.random.default_rng(2)
        for n in range(6):
            v = rng.integers(0, 2, size=lanes, dtype=U64)
            eng.set_pin_state(f"/l{n}_q0/out", v)
            ora.set_pin_state(f"/l{n}_q0/out", v)
        eng.run(33)
        ora.run(33)
        for p in eng.output_pins():
            assert np.array_equal(eng.get_pin_planes(p), ora.planes(p)), p
        n_rows = eng.program.n_persist + eng.program.n_scratch + 1
        plane = eng.rt.to_host_u64(eng._plane)
        rows = plane[:n_rows * eng.stride].reshape(n_rows, eng.stride)
        assert not rows[:, eng.words:].any(), "a kernel wrote into the padding columns of a row"
        assert not plane[n_rows * eng.stride:].any(), "a kernel wrote behind the last row"
    finally:
        if tune:
            eng.rt.set_tuning(tune[0], old)
