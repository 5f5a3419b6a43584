"""-m gpu: the CUDA engine (through the C ABI) against the golden vectors harvested from the
reference JIT and against the oracle on seeded inputs.  Bit-exact: integer work."""
import numpy as np
import pytest

import mcircuit_b200.core.descriptors as MD
from _helpers import (build, compare_engine_to_oracle, golden, lane_trick_planes,
                      run_register_case, run_trace_case)
from mcircuit_b200 import circuits, compiler
from mcircuit_b200.engine import B200Engine
from oracle import restate

pytestmark = pytest.mark.gpu
TRACES = golden("traces.json")


def gpu_factory(mode="resident", keep="all", map_pins=True, **kw):
    def make(root, burst_size, lanes=64):
        return B200Engine(root, burst_size, map_pins, lanes=lanes, mode=mode, keep=keep, **kw)
    return make


def test_native_library_is_the_one_running(cuda_rt):
    assert cuda_rt.cc[0] == 10, "expected a B200 (sm_100)"
    assert cuda_rt.sm_count >= 100
    import os
    maps = open("/proc/self/maps").read()
    assert "libmcb200.so" in maps


@pytest.mark.parametrize("case", TRACES, ids=[c["name"] for c in TRACES])
@pytest.mark.parametrize("mode,keep,map_pins", [("resident", "all", True), ("levels", "ports", True),
                                                ("levels", "all", False), ("resident", "ports", False),
                                                ("persistent", "ports", True), ("persistent", "all", False),
                                                ("levels_grouped", "ports", True)])
def test_gpu_traces(case, mode, keep, map_pins):
    run_trace_case(case, gpu_factory(mode, keep, map_pins), skip_unobservable=(keep == "ports"))


@pytest.mark.parametrize("back_edges", ["order", "latch"])
def test_gpu_traces_back_edge_policies(back_edges):
    for case in TRACES:
        for mode in ("levels", "resident"):
            run_trace_case(case, gpu_factory(mode, "all", True, back_edges=back_edges))


def test_gpu_traces_without_tma_and_smem(cuda_rt):
    """The resident kernel's four variants (TMA record staging on/off, smem state on/off)."""
    try:
        for tma in (0, 1):
            for smem in (0, 1):
                cuda_rt.set_tuning("resident_tma", tma)
                cuda_rt.set_tuning("resident_smem", smem)
                for case in TRACES[:12]:
                    run_trace_case(case, gpu_factory("resident", "all", True))
    finally:
        cuda_rt.set_tuning("resident_tma", 1)
        cuda_rt.set_tuning("resident_smem", 1)


@pytest.mark.parametrize("case", golden("register_intent.json"), ids=lambda c: c["name"])
@pytest.mark.parametrize("mode", ["resident", "levels"])
def test_gpu_register_intent(case, mode):
    run_register_case(case, gpu_factory(mode))


@pytest.mark.parametrize("case", golden("adder.json"), ids=lambda c: c["name"])
def test_gpu_adder(case):
    w = case["width"]
    top = MD.Composite()
    top.add_child("add", MD.Adder(w))
    vecs = case["vectors"]
    lanes = (len(vecs) + 63) // 64 * 64
    pad = lanes - len(vecs)
    eng = B200Engine(top, 1, True, lanes=lanes)
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
@pytest.mark.parametrize("mode,tile_words", [("levels", None), ("levels", 16), ("resident", None),
                                             ("resident", 32), ("persistent", None), ("persistent", 16)])
def test_gpu_lane_trick(case, mode, tile_words):
    ins, outs = lane_trick_planes(case)
    nw = case["n_words"]
    reps = 1 if tile_words is None else 9
    eng = B200Engine(build(case), 1, True, lanes=64 * nw * reps, mode=mode, keep="ports",
                     tile_words=tile_words)
    for p, words in ins.items():
        eng.set_pin_planes(p, np.tile(words, reps)[None, :])
    eng.step()
    for p, words in outs.items():
        assert np.array_equal(eng.get_pin_planes(p)[0], np.tile(words, reps)), p


@pytest.mark.parametrize("seed", range(200, 240))
def test_gpu_vs_oracle_all_lanes(seed):
    root, probes = circuits.random_circuit(MD, seed, n_nodes=22)
    d = compiler.flatten(root)
    lane_inputs = [(p, d.pin_width[d.find_pin(p)]) for p in ("/in0/pin", "/in1/pin")]
    mode = ("resident", "levels")[seed % 2]
    lanes = (64, 192, 1024, 4160)[seed % 4]
    compare_engine_to_oracle(lambda M: circuits.random_circuit(M, seed, n_nodes=22),
                             lambda r, bs, n: gpu_factory(mode)(r, bs, n),
                             lanes=lanes, steps=6, probes=probes, lane_inputs=lane_inputs, seed=seed)


@pytest.mark.parametrize("unroll", [1, 2, 4, 8])
@pytest.mark.parametrize("cache", [0, 1])
def test_gpu_level_kernel_variants(cuda_rt, unroll, cache):
    old_unroll = cuda_rt.set_tuning("level_unroll", unroll)      # returns the previous value
    old_cache = cuda_rt.set_tuning("level_cache", cache)
    try:
        case = golden("lane_trick.json")[3]
        ins, outs = lane_trick_planes(case)
        nw = case["n_words"]
        eng = B200Engine(build(case), 1, True, lanes=64 * nw * 7, mode="levels", keep="ports", tile_words=16)
        for p, words in ins.items():
            eng.set_pin_planes(p, np.tile(words, 7)[None, :])
        eng.step()
        for p, words in outs.items():
            assert np.array_equal(eng.get_pin_planes(p)[0], np.tile(words, 7)), p
    finally:
        # back to whatever was shipped (NOT a literal: the defaults live in mcb200.cu's Tuning)
        cuda_rt.set_tuning("level_unroll", old_unroll)
        cuda_rt.set_tuning("level_cache", old_cache)


def test_gpu_pack_unpack_stimulus_signature():
    root = circuits.ripple_adder(MD, 8)
    eng = B200Engine(root, 1, True, lanes=64 * 37)
    rng = np.random.default_rng(1)
    v = rng.integers(0, 2, size=eng.lanes, dtype=np.uint64)
    eng.set_pin_state("/a3/pin", v)
    assert np.array_equal(eng.get_pin_state("/a3/pin", lanes=True), v)
    assert np.array_equal(eng.get_pin_planes("/a3/pin")[0], restate.pack_lanes(v, 1)[0])
    n_bits = eng.randomize_inputs()
    assert n_bits == 17
    ins = eng.input_pins()
    for s, p in enumerate(ins):
        assert np.array_equal(eng.get_pin_planes(p)[0], restate.stimulus_word(s, np.arange(eng.words))), p
    eng.step()
    pc, cs = eng.signature()
    planes = np.stack([eng.get_pin_planes(p)[0] for p in eng.output_pins()])
    want_pc = np.array([sum(bin(int(w)).count("1") for w in row) for row in planes], dtype=np.uint64)
    cols = np.arange(eng.words, dtype=np.uint64)
    with np.errstate(over="ignore"):
        want_cs = (planes * (np.uint64(2) * cols + np.uint64(1))[None, :]).sum(axis=1, dtype=np.uint64)
    assert np.array_equal(pc, want_pc) and np.array_equal(cs, want_cs)
    # the adder really adds, on every lane
    a = sum(eng.get_pin_state(f"/a{i}/pin", lanes=True) << np.uint64(i) for i in range(8))
    b = sum(eng.get_pin_state(f"/b{i}/pin", lanes=True) << np.uint64(i) for i in range(8))
    s = sum(eng.get_pin_state(f"/s{i}/pin", lanes=True) << np.uint64(i) for i in range(8))
    s = s + (eng.get_pin_state("/cout/pin", lanes=True) << np.uint64(8))
    assert np.array_equal(s, a + b + eng.get_pin_state("/cin/pin", lanes=True))


def test_gpu_wide_pack_unpack():
    top = MD.Composite()
    top.add_child("n", MD.Not(64))
    eng = B200Engine(top, 1, True, lanes=64 * 5)
    rng = np.random.default_rng(2)
    v = rng.integers(0, 2**63, size=eng.lanes, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    eng.set_pin_state("/n/in", v)
    eng.step()
    assert np.array_equal(eng.get_pin_state("/n/out", lanes=True), ~v)
    assert eng.get_pin_state("/n/out") == int(~v[0])


def test_gpu_toggle_counts():
    eng = B200Engine(circuits.counter_circuit(MD), 1, True, lanes=128)
    eng.enable_toggle_counts(["/clk/out", "/r/out"])
    eng.run(8)
    t = eng.toggle_counts()
    assert t[0] == 8 * 128 and t[1] == 4 * 128 and not t[2:].any()


def test_gpu_burst_semantics():
    eng = B200Engine(circuits.counter_circuit(MD), 501, True)
    for _ in range(8):
        eng.step()
    eng.burst()
    assert (eng.get_pin_state("/clk/out"), eng.get_pin_state("/r/out"), eng.get_pin_state("/r/prevclock")) == (0, 1, 0)
    with pytest.raises(KeyError):
        eng.get_pin_state("/r/nope")


def test_gpu_run_host_pipelined_matches_resident():
    """Host buffers in / out through the tile-pipelined path == device-resident run."""
    import torch
    root = circuits.random_dag(MD, n_gates=3000, depth=15, n_inputs=48, seed=9)
    lanes = 64 * 200
    a = B200Engine(root, 1, True, lanes=lanes, mode="levels", keep="ports", tile_words=32)
    b = B200Engine(root, 1, True, lanes=lanes, mode="levels", keep="ports", tile_words=32)
    assert a.n_tiles == 7
    a.randomize_inputs()
    a.step()
    in_plan, out_plan = b.io_plan(b.input_pins()), b.io_plan(b.output_pins())
    h_in = torch.empty((in_plan["n"], b.words), dtype=torch.int64, pin_memory=True)
    h_out = torch.zeros((out_plan["n"], b.words), dtype=torch.int64, pin_memory=True)
    stim = restate.stimulus_word(np.arange(48)[:, None], np.arange(b.words)[None, :])
    h_in.numpy()[:] = stim.view(np.int64)
    for chunk_words in (4096, 64, 32):                 # one I/O chunk, chunks of 2 tiles, of 1 tile
        h_out.zero_()
        b.run_host(in_plan, h_in, out_plan, h_out, io_chunk_words=chunk_words)
        torch.cuda.synchronize()
        for k, p in enumerate(b.output_pins()):
            assert np.array_equal(h_out[k].numpy().view(np.uint64), a.get_pin_planes(p)[0]), (chunk_words, p)
    pa, ca = a.signature()
    pb, cb = b.signature()
    assert np.array_equal(pa, pb) and np.array_equal(ca, cb)


def test_reference_examples_run_unmodified(tmp_path):
    """The three example scripts of the reference, text unchanged except that they are
    restated here (the GPU box has no /root/reference): same imports, same calls."""
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "clock_example.py"
    script.write_text(
        "from core.descriptors import Composite, Not\n"
        "from core.simulator import JIT\n"
        "s = Composite()\nnot_ = Not()\n"
        "s.add_child('not', not_)\ns.connect('not', 'out', 'not', 'in')\n"
        "sim = JIT(s, 501, True)\n"
        "for i in range(10):\n    sim.step()\n    print(sim.get_pin_state('/not/out'))\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(root, "mcircuit_b200"))
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["1", "0", "1", "0", "1", "0", "1", "0", "1", "0"]


@pytest.mark.parametrize("unroll,hints", [(1, 0), (2, 1), (4, 1), (2, 0)])
def test_gpu_persistent_level_kernel_variants(cuda_rt, unroll, hints):
    try:
        cuda_rt.set_tuning("persist_unroll", unroll)
        cuda_rt.set_tuning("persist_hints", hints)
        root = circuits.random_dag(MD, n_gates=20_000, depth=40, n_inputs=256, seed=77)
        lanes = 64 * 320
        pa = ca = None
        for mode, kw in (("levels", {}), ("levels_grouped", {"tile_words": 64}),
                         ("persistent", {"tile_words": 64, "in_flight": 1}),
                         ("persistent", {"tile_words": 64, "in_flight": 2}),
                         ("persistent", {"tile_words": 64, "in_flight": 3}),
                         ("persistent", {"tile_words": 32, "in_flight": 4})):
            e = B200Engine(root, 1, True, lanes=lanes, keep="ports", mode=mode, **kw)
            e.randomize_inputs()
            e.run(1)
            p_, c_ = e.signature()
            if pa is None:
                pa, ca = p_, c_
            assert np.array_equal(pa, p_) and np.array_equal(ca, c_), (mode, kw)
        # sequential circuit, several steps inside one cooperative launch
        c = B200Engine(circuits.counter_circuit(MD), 1, True, lanes=128, mode="persistent")
        c.run(7)
        assert (c.get_pin_state("/clk/out"), c.get_pin_state("/r/out")) == (1, 0)
    finally:
        cuda_rt.set_tuning("persist_unroll", 2)
        cuda_rt.set_tuning("persist_hints", 1)


@pytest.mark.parametrize("word_bytes", [1, 2, 4, 8])
@pytest.mark.parametrize("batch", [4, 8])
def test_gpu_resident_word_and_batch_variants(word_bytes, batch):
    """The resident kernel with a lane word split among 8/4/2/1 threads and 4/8-record
    homogeneous batches: same results as the oracle."""
    for case in TRACES[:10] + TRACES[-6:]:
        run_trace_case(case, gpu_factory("resident", "all", True, resident_word_bytes=word_bytes,
                                         resident_batch=batch))
    root, probes = circuits.random_circuit(MD, 777, n_nodes=30)
    d = compiler.flatten(root)
    lane_inputs = [(p, d.pin_width[d.find_pin(p)]) for p in ("/in0/pin", "/in1/pin")]
    compare_engine_to_oracle(lambda M: circuits.random_circuit(M, 777, n_nodes=30),
                             lambda r, bs, n: gpu_factory("resident", "all", True, resident_word_bytes=word_bytes,
                                                          resident_batch=batch)(r, bs, n),
                             lanes=64 * 5, steps=9, probes=probes,
                             lane_inputs=lane_inputs, seed=5)


def test_gpu_capture_waveform_and_program_cache(tmp_path):
    from mcircuit_b200 import netlist_io, waveform
    eng = B200Engine(circuits.counter_circuit(MD), 1, True, lanes=128)
    pins = ["/clk/out", "/r/out"]
    samples = eng.capture(pins, 8)
    clk, r = waveform.lane_values(samples, eng.pin_widths(pins), lane=70)
    assert clk == [1, 0, 1, 0, 1, 0, 1, 0] and r == [1, 1, 0, 0, 1, 1, 0, 0]
    root = circuits.random_dag(MD, n_gates=800, depth=16, n_inputs=32, seed=2)
    prog = compiler.compile_circuit(root, keep="ports")
    netlist_io.save_program(prog, tmp_path / "p.npz")
    cached = netlist_io.load_program(tmp_path / "p.npz")
    a = B200Engine(root, 1, True, lanes=64 * 40, keep="ports", mode="levels", program=prog)
    b = B200Engine(None, 1, True, lanes=64 * 40, mode="levels", program=cached)
    for e in (a, b):
        e.randomize_inputs()
        e.step()
    pa, ca = a.signature()
    pb, cb = b.signature()
    assert np.array_equal(pa, pb) and np.array_equal(ca, cb)


@pytest.mark.parametrize("streams", [2, 3])
def test_gpu_tiles_on_several_streams(streams):
    """Independent lane tiles alternating over several CUDA streams (own scratch plane each)."""
    import torch
    root = circuits.random_dag(MD, n_gates=6000, depth=20, n_inputs=64, seed=21)
    lanes = 64 * 32 * 7
    a = B200Engine(root, 1, True, lanes=lanes, keep="ports", mode="levels", tile_words=32)
    b = B200Engine(root, 1, True, lanes=lanes, keep="ports", mode="levels", tile_words=32, streams=streams)
    assert b.n_tiles == 7 and b.n_streams == streams
    for e in (a, b):
        e.randomize_inputs()
        e.run(2)
    pa, ca = a.signature()
    pb, cb = b.signature()
    assert np.array_equal(pa, pb) and np.array_equal(ca, cb)
    in_plan, out_plan = b.io_plan(b.input_pins()), b.io_plan(b.output_pins())
    h_in = torch.empty((in_plan["n"], b.words), dtype=torch.int64, pin_memory=True)
    h_out = torch.zeros((out_plan["n"], b.words), dtype=torch.int64, pin_memory=True)
    h_in.numpy()[:] = restate.stimulus_word(np.arange(64)[:, None], np.arange(b.words)[None, :]).view(np.int64)
    for chunk_words in (4096, 96):
        h_out.zero_()
        b.run_host(in_plan, h_in, out_plan, h_out, io_chunk_words=chunk_words)
        torch.cuda.synchronize()
        for k, p in enumerate(b.output_pins()):
            assert np.array_equal(h_out[k].numpy().view(np.uint64), a.get_pin_planes(p)[0]), (chunk_words, p)


def test_reference_side_stub_of_integration_md():
    """examples/reference_side_stub.py (the binding INTEGRATION.md shows) against the goldens."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("reference_side_stub",
                                                  os.path.join(root, "examples", "reference_side_stub.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for case in TRACES[:8] + TRACES[-4:]:
        run_trace_case(case, lambda r, bs: mod.B200(r, bs, True))


def test_gpu_edge_cases_empty_ragged_wide():
    eng = B200Engine(MD.Composite(), 3, True)
    eng.step()
    eng.burst()
    assert eng.steps_done == 5
    top = MD.Composite()
    top.add_child("g", MD.Gate(MD.Gate.XOR, 64, 64, True))
    lanes = 64 * 3
    eng = B200Engine(top, 1, True, lanes=lanes)
    ora = restate.RestatedJIT(top, 1, True, lanes=lanes)
    rng = np.random.default_rng(9)
    for i in range(64):
        v = rng.integers(0, 2**63, size=lanes, dtype=np.uint64) * np.uint64(2) + np.uint64(i & 1)
        eng.set_pin_state(f"/g/in{i}", v)
        ora.set_pin_state(f"/g/in{i}", v)
    eng.step()
    ora.step()
    assert np.array_equal(eng.get_pin_state("/g/out", lanes=True), ora.get_pin_state("/g/out", lanes=True))
    # a single lane word, an odd number of lane words, and a tile boundary that is not a multiple of the tile
    root = circuits.random_dag(MD, n_gates=900, depth=9, n_inputs=20, seed=31)
    ref = None
    for lanes_, kw in ((64, {}), (64 * 37, {}), (64 * 37, {"tile_words": 16, "mode": "levels"}),
                       (64 * 37, {"tile_words": 16, "mode": "persistent"}), (64 * 37, {"mode": "resident"})):
        e = B200Engine(root, 1, True, lanes=lanes_, keep="ports", **kw)
        e.randomize_inputs()
        e.step()
        planes = np.stack([e.get_pin_planes(p)[0] for p in e.output_pins()])
        if lanes_ == 64 * 37:
            if ref is None:
                ref = planes
            assert np.array_equal(planes, ref), kw
        else:
            assert ref is None or np.array_equal(planes[:, :1], ref[:, :1])


def test_gpu_c_abi_entry_points_directly(cuda_rt):
    """mcb_eval_levels, mcb_latch, mcb_fill_rows, mcb_capture_rows and mcb_copy2d called the way
    a foreign host would: raw device pointers + sizes, no engine in between."""
    import torch
    rt = cuda_rt
    # netlist: level 0: s2 = s0 & s1, s3 = s0 ^ s1 ; level 1: s4 = ~(s2 | s3) ; latch: s5 = s4
    recs = np.array([[compiler.OP_AND, 0, 1, 2], [compiler.OP_XOR, 0, 1, 3],
                     [compiler.OP_OR | compiler.OP_NEG, 2, 3, 4]], dtype=np.uint32)
    latch = np.array([[compiler.OP_BUF, 4, 0, 5]], dtype=np.uint32)
    off = np.array([0, 2, 3], dtype=np.int32)
    words, stride, n_slots = 37, 48, 6
    plane = rt.zeros_words(n_slots * stride)
    st = rt.make_state(plane, stride, None, 0, n_slots, n_slots, words)
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2**63, size=words, dtype=np.uint64)
    b = rng.integers(0, 2**63, size=words, dtype=np.uint64)
    h = torch.zeros((2, words), dtype=torch.int64).pin_memory()
    h.numpy()[0] = a.view(np.int64)
    h.numpy()[1] = b.view(np.int64)
    rt.copy2d(plane.data_ptr(), stride * 8, h.data_ptr(), words * 8, words * 8, 2, True)        # H2D rows 0, 1
    d_recs, d_latch = rt.to_device(recs), rt.to_device(latch)
    rt.eval_levels(d_recs, off, 0, 2, st)
    rt.latch(d_latch, 1, st)
    out = torch.zeros((2, words), dtype=torch.int64).pin_memory()
    rt.copy2d(out.data_ptr(), words * 8, plane.data_ptr() + 4 * stride * 8, stride * 8, words * 8, 2, False)
    cap = rt.zeros_words(2 * 64)
    rt.capture_rows(rt.to_device(np.array([2, 3], dtype=np.int32)), 2, st, cap, 0, 64)
    rt.synchronize()
    want = ~((a & b) | (a ^ b))
    assert np.array_equal(out.numpy().view(np.uint64)[0], want)               # slot 4
    assert np.array_equal(out.numpy().view(np.uint64)[1], want)               # slot 5, latched copy
    c = rt.to_host_u64(cap).reshape(2, 64)[:, :words]
    assert np.array_equal(c[0], a & b) and np.array_equal(c[1], a ^ b)
    # argument validation reaches the caller as an exception with the library's message
    bad = rt.make_state(plane, stride - 1, None, 0, n_slots, n_slots, words)    # odd stride
    with pytest.raises(RuntimeError, match="strides must be even"):
        rt.eval_levels(d_recs, off, 0, 2, bad)


# ---- K1t: the level kernel with rows staged through shared memory by TMA bulk copies --------------------------
@pytest.fixture
def level_tma(cuda_rt):
    old = cuda_rt.set_tuning("level_tma", 1)
    yield
    cuda_rt.set_tuning("level_tma", old)


@pytest.mark.parametrize("case", [c for c in golden("traces.json") if not c["name"].startswith("fuzz_seed3")][:24],
                         ids=lambda c: c["name"])
def test_gpu_level_tma_traces(level_tma, case):
    run_trace_case(case, lambda root, bs: B200Engine(root, bs, True, mode="levels"))


@pytest.mark.parametrize("case", golden("lane_trick.json"), ids=lambda c: c["name"])
@pytest.mark.parametrize("tile_words", [None, 16, 48])
def test_gpu_level_tma_lane_trick(level_tma, case, tile_words):
    """Reference-JIT words on ragged lane counts and tile sizes (rows of 8 x odd words, several tiles, 2 streams)."""
    ins, outs = lane_trick_planes(case)
    nw = case["n_words"]
    reps = 11
    eng = B200Engine(build(case), 1, True, lanes=64 * nw * reps, mode="levels", keep="ports", tile_words=tile_words)
    for p, words in ins.items():
        eng.set_pin_planes(p, np.tile(words, reps)[None, :])
    eng.step()
    for p, words in outs.items():
        assert np.array_equal(eng.get_pin_planes(p)[0], np.tile(words, reps)), p


@pytest.mark.parametrize("seed", range(240, 260))
def test_gpu_level_tma_vs_oracle_all_lanes(level_tma, seed):
    root, probes = circuits.random_circuit(MD, seed, n_nodes=22)
    d = compiler.flatten(root)
    lane_inputs = [(p, d.pin_width[d.find_pin(p)]) for p in ("/in0/pin", "/in1/pin")]
    lanes = (64, 192, 1024, 4160, 64 * 513)[seed % 5]
    compare_engine_to_oracle(lambda M: circuits.random_circuit(M, seed, n_nodes=22),
                             lambda r, bs, n: gpu_factory("levels")(r, bs, n),
                             lanes=lanes, steps=6, probes=probes, lane_inputs=lane_inputs, seed=seed)


def test_gpu_level_tma_c5_20k_vs_oracle(level_tma):
    from oracle import cref, restate
    root = circuits.random_dag(MD, n_gates=20_000, depth=40, n_inputs=256, seed=77)
    lanes = 64 * 1100                                    # 512-word tiles, the last one ragged (76 words)
    eng = B200Engine(root, 1, True, lanes=lanes, keep="ports", mode="levels", tile_words=512)
    assert eng.n_tiles == 3
    eng.randomize_inputs()
    eng.step()
    eng.step()
    ora = cref.CRef(root, 1, True, lanes=lanes)
    ora.set_rows(ora._base("/i0/pin"), restate.stimulus_word(np.arange(256)[:, None], np.arange(eng.words)[None, :]))
    ora.step()
    want = ora.rows(ora._base("/o0/pin"), 500)
    got = np.stack([eng.get_pin_planes(p)[0] for p in eng.output_pins()])
    assert np.array_equal(got, want)
