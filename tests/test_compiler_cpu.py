"""not-gpu: the netlist compiler + engine host logic, executed by the numpy test double
(tests/_cpu_runtime.py) and checked against the oracle and the golden vectors."""
import numpy as np
import pytest

import mcircuit_b200.core.descriptors as MD
from _cpu_runtime import CpuRuntime
from _helpers import (build, compare_engine_to_oracle, golden, lane_trick_planes,
                      run_register_case, run_trace_case)
from mcircuit_b200 import circuits, compiler
from mcircuit_b200.engine import B200Engine

HEAVY = ("c4_gate_level_lfsr32",)          # 97 probes x 400 steps: in full on the GPU and on the oracles, shortened on the numpy test double
ALL_TRACES = golden("traces.json")
TRACES = [c for c in ALL_TRACES if c["name"] not in HEAVY]


def engine_factory(mode="resident", keep="all", map_pins=True, **kw):
    def make(root, burst_size, lanes=64):
        return B200Engine(root, burst_size, map_pins, lanes=lanes, runtime=CpuRuntime(),
                          mode=mode, keep=keep, **kw)
    return make


@pytest.mark.parametrize("case", TRACES, ids=[c["name"] for c in TRACES])
@pytest.mark.parametrize("mode,keep,map_pins", [("resident", "all", True), ("levels", "ports", True),
                                                ("levels", "all", False), ("persistent", "ports", True)])
def test_engine_traces(case, mode, keep, map_pins):
    run_trace_case(case, engine_factory(mode, keep, map_pins), skip_unobservable=(keep == "ports"))


@pytest.mark.parametrize("case", TRACES, ids=[c["name"] for c in TRACES])
@pytest.mark.parametrize("back_edges", ["order", "latch"])
def test_engine_traces_back_edge_policies(case, back_edges):
    """Reads of last step's value: settled by ordering or by shadow slot + latch level."""
    run_trace_case(case, engine_factory("levels", "all", True, back_edges=back_edges))


def test_back_edge_policies_shape():
    root, _ = circuits.random_circuit(MD, 0, n_nodes=20)
    order = compiler.compile_circuit(root, back_edges="order")
    latch = compiler.compile_circuit(root, back_edges="latch")
    assert not order.has_latch and not order.sig_shadow
    assert latch.has_latch and latch.sig_shadow and latch.n_ops == order.n_ops + len(latch.sig_shadow)
    assert latch.n_gates == order.n_gates
    c4 = compiler.compile_circuit(circuits.lfsr_pipeline(MD, 4, 8, 6, taps=(7, 5, 1, 0)))
    assert not c4.sig_shadow          # prevclock read-then-write is settled by ordering


@pytest.mark.parametrize("case", golden("register_intent.json"), ids=lambda c: c["name"])
def test_engine_register_intent(case):
    run_register_case(case, engine_factory("levels"))


@pytest.mark.parametrize("case", golden("adder.json"), ids=lambda c: c["name"])
def test_engine_adder(case):
    w = case["width"]
    top = MD.Composite()
    top.add_child("add", MD.Adder(w))
    vecs = case["vectors"]
    lanes = (len(vecs) + 63) // 64 * 64
    pad = lanes - len(vecs)
    eng = B200Engine(top, 1, True, lanes=lanes, runtime=CpuRuntime())
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
@pytest.mark.parametrize("tile_words", [None, 16])
def test_engine_lane_trick(case, tile_words):
    ins, outs = lane_trick_planes(case)
    nw = case["n_words"]
    # replicate the words so that several tiles exist when tiling is forced
    reps = 1 if tile_words is None else max(5, -(-40 // nw))
    eng = B200Engine(build(case), 1, True, lanes=64 * nw * reps, runtime=CpuRuntime(), mode="levels",
                     keep="ports", tile_words=tile_words)
    if tile_words is not None:
        assert eng.n_tiles > 1
    for p, words in ins.items():
        eng.set_pin_planes(p, np.tile(words, reps)[None, :])
    eng.step()
    for p, words in outs.items():
        assert np.array_equal(eng.get_pin_planes(p)[0], np.tile(words, reps)), p


@pytest.mark.parametrize("seed", range(100, 130))
def test_engine_vs_oracle_all_lanes(seed):
    root, probes = circuits.random_circuit(MD, seed, n_nodes=22)
    d = compiler.flatten(root)
    lane_inputs = [("/in0/pin", d.pin_width[d.find_pin("/in0/pin")]),
                   ("/in1/pin", d.pin_width[d.find_pin("/in1/pin")])]
    mode = ("resident", "levels")[seed % 2]
    compare_engine_to_oracle(lambda M: circuits.random_circuit(M, seed, n_nodes=22),
                             lambda r, bs, lanes: engine_factory(mode)(r, bs, lanes),
                             lanes=128, steps=6, probes=probes, lane_inputs=lane_inputs, seed=seed)


def test_lfsr_pipeline_small_vs_oracle():
    probes = [f"/tap{n}/pin" for n in range(3)] + ["/pipe_out/pin"] + [f"/l1_q{b}/out" for b in range(8)]
    pokes = [(f"/l{n}_q{b}/out", (n + b) & 1) for n in range(3) for b in range(8)]
    for mode in ("resident", "levels"):
        compare_engine_to_oracle(lambda M: circuits.lfsr_pipeline(M, 3, 8, 10, taps=(7, 5, 1, 0)),
                                 lambda r, bs, lanes: engine_factory(mode)(r, bs, lanes),
                                 lanes=64, steps=40, probes=probes, pokes=pokes)


def test_burst_and_run_and_tick():
    eng = engine_factory()(circuits.counter_circuit(MD), 501)
    for _ in range(8):
        eng.step()
    eng.burst()
    assert eng.steps_done == 8 + 502
    assert (eng.get_pin_state("/clk/out"), eng.get_pin_state("/r/out"), eng.get_pin_state("/r/prevclock")) == (0, 1, 0)
    eng.tick()
    assert eng.steps_done == 8 + 502 + 2


def test_constants_visible_before_first_step_and_masking():
    top = MD.Composite()
    top.add_child("c", MD.Constant(5, 0b10110))
    top.add_child("n", MD.Not(5))
    top.connect("c", "out", "n", "in")
    eng = engine_factory()(top, 1)
    assert eng.get_pin_state("/c/out") == 0b10110
    assert eng.get_pin_state("/n/in") == 0b10110          # aliases its source (Q7)
    eng.step()
    assert eng.get_pin_state("/n/out") == 0b01001
    eng.set_pin_state("/n/in", 0xFFFF)                     # written through to the source, masked (Q9)
    assert eng.get_pin_state("/c/out") == 0b11111


def test_unknown_pin_raises_keyerror():
    eng = engine_factory()(circuits.counter_circuit(MD), 1)
    for bad in ("/nope/out", "/clk/nope", "clk/out", "/clk", "/clk/out/x"):
        with pytest.raises(KeyError):
            eng.get_pin_state(bad)
    with pytest.raises(KeyError):
        eng.set_pin_state("/r/nope", 1)


def test_build_time_validation():
    top = MD.Composite()
    top.add_child("a", MD.Not(8))
    top.add_child("b", MD.Not(4))
    top.connect("a", "out", "b", "in")
    with pytest.raises(compiler.CircuitError):
        compiler.compile_circuit(top)
    top = MD.Composite()
    top.add_child("a", MD.Not())
    top.add_child("b", MD.Not())
    top.add_child("c", MD.Not())
    top.connect("a", "out", "c", "in")
    top.connect("b", "out", "c", "in")
    with pytest.raises(compiler.CircuitError):
        compiler.compile_circuit(top)
    top = MD.Composite()
    top.add_child("a", MD.Not())
    top.add_child("b", MD.Not())
    top.connect("a", "out", "b", "nope")
    with pytest.raises(KeyError):
        compiler.compile_circuit(top)
    top = MD.Composite()
    top.add_child("g", MD.Gate(MD.Gate.AND, 1, 0))
    with pytest.raises(compiler.CircuitError):
        compiler.compile_circuit(top)


def test_ports_mode_recycles_and_reports_unobservable():
    root = circuits.random_dag(MD, n_gates=600, depth=12, n_inputs=16, seed=5)
    full = compiler.compile_circuit(root, keep="all")
    lean = compiler.compile_circuit(root, keep="ports")
    assert full.n_scratch == 0 and lean.n_scratch > 0
    assert lean.n_persist == 16 + 50                       # inputs + last-level outputs
    assert lean.n_slots < full.n_slots / 2
    assert lean.n_gates == full.n_gates == 600
    assert lean.n_levels == 12 and not lean.has_latch
    assert lean.gate_words_bytes == 600 * 24               # 24 B per 2-input gate word
    eng = B200Engine(root, 1, True, lanes=64, runtime=CpuRuntime(), keep="ports")
    eng.step()
    with pytest.raises(compiler.PinNotObservable):
        eng.get_pin_state("/g3_7/out")
    eng2 = B200Engine(root, 1, True, lanes=64, runtime=CpuRuntime(), keep="ports", observe=["/g3_7/out"])
    eng2.step()
    eng2.get_pin_state("/g3_7/out")


def test_levelization_invariants():
    """Within a level no record reads what another writes; slot reuse never clobbers a value
    that is still to be read."""
    for seed in range(20):
        root, _ = circuits.random_circuit(MD, seed, n_nodes=25)
        for keep in ("all", "ports"):
            p = compiler.compile_circuit(root, keep=keep)
            rec, off = p.records.astype(np.int64), p.level_off
            for lv in range(p.n_levels):
                r = rec[off[lv]:off[lv + 1]]
                op = r[:, 0] & 0x7F
                reads = set()
                for k in range(len(r)):
                    ar = compiler.OP_ARITY[int(op[k])]
                    ins = [r[k, 1], r[k, 2], r[k, 0] >> 8][:ar]
                    reads.update((int(x), k) for x in ins)
                writes = {int(r[k, 3]): k for k in range(len(r))}
                assert len(writes) == len(r)
                for slot, k in reads:
                    assert slot not in writes or writes[slot] == k, (seed, keep, lv)


def test_signature_and_stimulus_and_toggles():
    root = circuits.ripple_adder(MD, 8)
    eng = B200Engine(root, 1, True, lanes=256, runtime=CpuRuntime())
    n_bits = eng.randomize_inputs()
    assert n_bits == 17
    eng.step()
    a = sum(eng.get_pin_state(f"/a{i}/pin", lanes=True) << np.uint64(i) for i in range(8))
    b = sum(eng.get_pin_state(f"/b{i}/pin", lanes=True) << np.uint64(i) for i in range(8))
    cin = eng.get_pin_state("/cin/pin", lanes=True)
    s = sum(eng.get_pin_state(f"/s{i}/pin", lanes=True) << np.uint64(i) for i in range(8))
    s = s + (eng.get_pin_state("/cout/pin", lanes=True) << np.uint64(8))
    assert np.array_equal(s, a + b + cin)
    pc, cs = eng.signature()
    planes = np.stack([eng.get_pin_planes(p)[0] for p in eng.output_pins()])
    want_pc = np.array([sum(bin(int(w)).count("1") for w in row) for row in planes], dtype=np.uint64)
    assert np.array_equal(pc, want_pc)
    from oracle.restate import stimulus_word
    assert np.array_equal(eng.get_pin_planes("/cin/pin")[0], stimulus_word(0, np.arange(4)))
    # toggles on the clocked counter: /clk/out flips every step in every lane
    eng = B200Engine(circuits.counter_circuit(MD), 1, True, lanes=128, runtime=CpuRuntime())
    eng.enable_toggle_counts(["/clk/out", "/r/out"])
    eng.run(8)
    t = eng.toggle_counts()
    assert t[0] == 8 * 128 and t[1] == 4 * 128 and not t[2:].any()


def test_descriptor_api_surface():
    g = MD.Gate(MD.Gate.XOR, 4, 3, True)
    assert list(g.all_inputs()) == [("in0", 4), ("in1", 4), ("in2", 4)]
    assert list(g.all_outputs()) == [("out", 4)] and list(g.all_internals()) == []
    assert list(MD.Register(3).all_pins()) == [("clock", 1), ("data", 3), ("out", 3), ("prevclock", 1)]
    assert list(MD.Counter(2).all_pins()) == [("clock", 1), ("out", 2), ("prevclock", 1)]
    assert list(MD.Clock().all_pins()) == [("out", 1), ("count", 64)]
    assert list(MD.Adder(5).all_pins()) == [("cin", 1), ("a", 5), ("b", 5), ("cout", 1), ("sum", 5)]
    assert list(MD.ExposedPin(MD.ExposedPin.IN, 2).all_outputs()) == [("pin", 2)]
    assert list(MD.ExposedPin(MD.ExposedPin.OUT, 2).all_inputs()) == [("pin", 2)]
    assert (MD.ExposedPin.IN, MD.ExposedPin.OUT, MD.Gate.AND, MD.Gate.OR, MD.Gate.XOR) == (0, 1, 0, 1, 2)
    c = MD.Composite()
    c.add_child("x", MD.ExposedPin(MD.ExposedPin.IN))
    sub = MD.Composite()
    sub.add_child("p", MD.ExposedPin(MD.ExposedPin.IN))
    c.add_child("s", sub)
    c.add_child("n", MD.Not())
    c.connect("x", "", "s", "p")
    c.connect("x", "whatever", "n", "in")
    assert ("x", "pin", "s/p", "pin") in c.connections and ("x", "pin", "n", "in") in c.connections
    assert list(c.graph.edges) == [("s", "x"), ("n", "x")]            # dst -> src
    assert c.graph.nodes["n"]["label"] == "n" and c.get_child("n") is c.graph.nodes["n"]["descriptor"]
    assert list(c.all_inputs()) == [("x", 1)] and list(c.all_outputs()) == []
    k = g.clone()
    assert k is not g and (k.op, k.width, k.num_inputs, k.negated) == (2, 4, 3, True)
    from mcircuit_b200.core import simulator
    assert issubclass(simulator.JIT, simulator.Executor)
    assert list(simulator.iter_simulation_pins(c))[0] == ("/x/pin", 1, "out")


def test_streaming_hints_shape():
    root = circuits.random_dag(MD, n_gates=600, depth=12, n_inputs=16, seed=5)
    p = compiler.compile_circuit(root, keep="ports")
    h = p.records_hinted
    assert np.array_equal(h[:, [0, 3]], p.records[:, [0, 3]])
    assert np.array_equal(h[:, 1:3] & 0x7FFFFFFF, p.records[:, 1:3])
    lv0 = slice(p.level_off[0], p.level_off[1])
    assert (h[lv0, 1] >> 31).all() and (h[lv0, 2] >> 31).all()          # level 0 reads inputs only
    lv5 = slice(p.level_off[5], p.level_off[6])
    assert not (h[lv5, 1] >> 31).any()                                   # in0 comes from level 4
    assert (h[lv5, 2] >> 31).mean() > 0.5                                # in1 mostly from far back


@pytest.mark.parametrize("batch", [4, 8])
def test_batch_schedule_hazards_and_results(batch):
    """schedule_batches: every record placed exactly once, batches are homogeneous in opcode,
    nothing inside a batch reads what the batch writes (except in place), and the engine
    (numpy double, loads-before-stores per batch) still matches the oracle."""
    root = circuits.lfsr_pipeline(MD, 3, 8, 10, taps=(7, 5, 1, 0))
    p = compiler.compile_circuit(root, keep="ports")
    trash = p.n_slots
    rec, nb = compiler.schedule_batches(p.records, p.level_off, trash, batch)
    assert len(rec) == nb * batch
    real = rec[rec[:, 3] != trash]
    assert len(real) == p.n_ops
    assert sorted(map(tuple, real.tolist())) == sorted(map(tuple, p.records.tolist()))
    for b in range(nb):
        grp = rec[b * batch:(b + 1) * batch].astype(np.int64)
        assert len(set((grp[:, 0] & 0x7F).tolist())) == 1                # one base opcode
        grp = grp[grp[:, 3] != trash]
        writes = {int(r[3]): k for k, r in enumerate(grp)}
        assert len(writes) == len(grp)
        for k, r in enumerate(grp):
            ar = compiler.OP_ARITY[int(r[0] & 0x7F)]
            for s_ in [r[1], r[2], r[0] >> 8][:ar]:
                assert int(s_) not in writes or writes[int(s_)] == k      # only an in-place read
    probes = [f"/tap{n}/pin" for n in range(3)] + ["/pipe_out/pin"]
    pokes = [(f"/l{n}_q{b}/out", (n + b) & 1) for n in range(3) for b in range(8)]
    compare_engine_to_oracle(lambda M: circuits.lfsr_pipeline(M, 3, 8, 10, taps=(7, 5, 1, 0)),
                             lambda r, bs, lanes: engine_factory("resident", "ports", resident_batch=batch)(r, bs, lanes),
                             lanes=64, steps=30, probes=probes, pokes=pokes)


def test_edge_cases_empty_ragged_wide():
    # empty circuit: nothing to evaluate, stepping is a no-op
    eng = engine_factory()(MD.Composite(), 3)
    eng.step()
    eng.burst()
    assert eng.steps_done == 5 and eng.program.n_ops == 0 and eng.gate_evals(7) == 0
    # ports only, wired straight through (pure aliasing, no primitive at all)
    top = MD.Composite()
    top.add_child("i", MD.ExposedPin(MD.ExposedPin.IN, 7))
    top.add_child("o", MD.ExposedPin(MD.ExposedPin.OUT, 7))
    top.connect("i", "", "o", "")
    eng = engine_factory()(top, 1)
    eng.set_pin_state("/i/pin", 0x55)
    eng.step()
    assert eng.get_pin_state("/o/pin") == 0x55 and eng.program.n_ops == 0
    # lanes must be whole 64-lane words
    for bad in (0, 1, 63, 100):
        with pytest.raises(ValueError):
            B200Engine(top, 1, True, lanes=bad, runtime=CpuRuntime())
    # the widest things the reference's GUI allows (editors.py:108-124): 64 inputs, 64 bits
    top = MD.Composite()
    top.add_child("g", MD.Gate(MD.Gate.XOR, 64, 64, True))
    eng = B200Engine(top, 1, True, lanes=64 * 3, runtime=CpuRuntime())
    ora = __import__("oracle.restate", fromlist=["x"]).RestatedJIT(top, 1, True, lanes=64 * 3)
    rng = np.random.default_rng(9)
    for i in range(64):
        v = rng.integers(0, 2**63, size=64 * 3, dtype=np.uint64) * np.uint64(2) + np.uint64(i & 1)
        eng.set_pin_state(f"/g/in{i}", v)
        ora.set_pin_state(f"/g/in{i}", v)
    eng.step()
    ora.step()
    assert np.array_equal(eng.get_pin_state("/g/out", lanes=True), ora.get_pin_state("/g/out", lanes=True))
    assert eng.program.n_gates == 64 * 32            # 63 two-input folds become 32 three-input ops per bit


def test_deep_hierarchy_and_shared_instances():
    """The same Composite object under many names and at several depths (instancing by path,
    examples/raw_counter.py:104-105)."""
    inv = MD.Composite()
    inv.add_child("a", MD.ExposedPin(MD.ExposedPin.IN))
    inv.add_child("n", MD.Not())
    inv.add_child("y", MD.ExposedPin(MD.ExposedPin.OUT))
    inv.connect("a", "", "n", "in")
    inv.connect("n", "out", "y", "")
    level = inv
    for depth in range(5):                          # wrap: two instances of the previous level in series
        w = MD.Composite()
        w.add_child("a", MD.ExposedPin(MD.ExposedPin.IN))
        w.add_child("u", level)
        w.add_child("v", level)
        w.add_child("y", MD.ExposedPin(MD.ExposedPin.OUT))
        w.connect("a", "", "u", "a")
        w.connect("u", "y", "v", "a")
        w.connect("v", "y", "y", "")
        level = w
    top = MD.Composite()
    top.add_child("x", MD.ExposedPin(MD.ExposedPin.IN))
    top.add_child("c", level)
    top.add_child("z", MD.ExposedPin(MD.ExposedPin.OUT))
    top.connect("x", "", "c", "a")
    top.connect("c", "y", "z", "")
    eng = engine_factory("levels", "all")(top, 1)
    assert eng.program.n_gates == 32 and eng.program.n_levels == 32      # 2^5 inverters in series
    for v in (0, 1):
        eng.set_pin_state("/x/pin", v)
        eng.step()
        assert eng.get_pin_state("/z/pin") == v                           # even number of inverters
        assert eng.get_pin_state("/c/u/u/u/u/u/n/out") == 1 - v           # the very first inverter
        assert eng.get_pin_state("/c/v/v/v/v/v/y/pin") == v


@pytest.mark.parametrize("mode", ["resident", "levels"])
def test_shared_edge_detectors_are_exact(mode):
    """share_edge_detectors=True: one rise / prevclock per clock group, same results on every
    pin (the members' prevclock pins view the leader's bit)."""
    args = dict(n_lfsr=3, lfsr_bits=8, stages=10, taps=(7, 5, 1, 0))
    plain = compiler.compile_circuit(circuits.lfsr_pipeline(MD, **args))
    shared = compiler.compile_circuit(circuits.lfsr_pipeline(MD, **args), share_edge_detectors=True)
    n_regs = 3 * 8 + 10
    assert shared.n_gates == plain.n_gates - 2 * (n_regs - 1)
    probes = ([f"/tap{n}/pin" for n in range(3)] + ["/pipe_out/pin", "/clk/out"] +
              [f"/l{n}_q{b}/prevclock" for n in range(3) for b in (0, 3, 7)] + [f"/p{k}/prevclock" for k in (0, 9)] +
              [f"/l1_q{b}/out" for b in range(8)])
    pokes = [(f"/l{n}_q{b}/out", (n + b) & 1) for n in range(3) for b in range(8)]
    compare_engine_to_oracle(lambda M: circuits.lfsr_pipeline(M, **args),
                             lambda r, bs, lanes: engine_factory(mode, share_edge_detectors=True)(r, bs, lanes),
                             lanes=64, steps=40, probes=probes, pokes=pokes)


@pytest.mark.parametrize("seed", range(400, 430))
def test_shared_edge_detectors_fuzz(seed):
    prims = ("Gate", "Not", "Constant", "Clock", "Counter", "Counter", "Adder", "Register", "Register")
    root, probes = circuits.random_circuit(MD, seed, n_nodes=26, primitives=prims)
    d = compiler.flatten(root)
    lane_inputs = [(p, d.pin_width[d.find_pin(p)]) for p in ("/in0/pin", "/in1/pin")]
    compare_engine_to_oracle(lambda M: circuits.random_circuit(M, seed, n_nodes=26, primitives=prims),
                             lambda r, bs, lanes: engine_factory(("resident", "levels")[seed % 2],
                                                                 share_edge_detectors=True)(r, bs, lanes),
                             lanes=64, steps=8, probes=probes, lane_inputs=lane_inputs, seed=seed)


def test_sort_level_by_fanin_keeps_levels_and_results():
    root = circuits.random_dag(MD, n_gates=1200, depth=12, n_inputs=32, seed=6)
    a = compiler.compile_circuit(root, keep="ports")
    b = compiler.compile_circuit(root, keep="ports", sort_level_by_fanin=True)
    assert np.array_equal(a.level_off, b.level_off) and a.n_slots == b.n_slots
    for lv in range(a.n_levels):
        ra = a.records[a.level_off[lv]:a.level_off[lv + 1]]
        rb = b.records[b.level_off[lv]:b.level_off[lv + 1]]
        assert sorted(map(tuple, ra.tolist())) == sorted(map(tuple, rb.tolist()))     # same records
        assert (np.diff(rb[:, 2].astype(np.int64)) >= 0).all()                         # ordered by in1 slot
    ea = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), mode="levels", program=a)
    eb = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), mode="resident", program=b)
    for e in (ea, eb):
        e.randomize_inputs()
        e.step()
    assert all(np.array_equal(ea.get_pin_planes(p), eb.get_pin_planes(p)) for p in ea.output_pins())


# ---- the Adder primitive lowered as a word-level parallel-prefix network ---------------------------
@pytest.mark.parametrize("case", golden("adder.json"), ids=lambda c: c["name"])
@pytest.mark.parametrize("mode", ["levels", "serial"])
def test_adder_prefix_lowering_matches_reference_jit_vectors(case, mode):
    """Same function as the ripple lowering (and as the reference JIT: simulator.py:135-147), logarithmic depth."""
    w = case["width"]
    top = MD.Composite()
    top.add_child("add", MD.Adder(w))
    vecs = case["vectors"]
    lanes = (len(vecs) + 63) // 64 * 64
    pad = lanes - len(vecs)
    eng = B200Engine(top, 1, True, lanes=lanes, runtime=CpuRuntime(), mode=mode, adder_prefix=True)
    col = lambda k: np.array([v[k] for v in vecs] + [0] * pad, dtype=np.uint64)
    eng.set_pin_state("/add/a", col(0))
    eng.set_pin_state("/add/b", col(1))
    eng.set_pin_state("/add/cin", col(2))
    eng.step()
    s = eng.get_pin_state("/add/sum", lanes=True)
    c = eng.get_pin_state("/add/cout", lanes=True)
    for i, (ws, wc) in enumerate(case["results"]):
        assert (int(s[i]), int(c[i])) == (ws, wc), (w, vecs[i])
    if mode == "levels" and w >= 16:
        ripple = compiler.compile_circuit(top)
        assert eng.program.n_levels <= 2 * int(np.ceil(np.log2(w))) + 8 < ripple.n_levels


def test_adder_prefix_lowering_in_random_circuits_with_feedback():
    """Adders whose outputs feed their own inputs (aliasing) and sit in feedback loops: prefix == ripple == oracle."""
    from oracle import restate
    for seed in range(500, 512):
        prims = ("Gate", "Not", "Constant", "Clock", "Counter", "Adder", "Adder", "Adder", "Register")
        root, probes = circuits.random_circuit(MD, seed, n_nodes=22, primitives=prims)
        ora = restate.RestatedJIT(root, 3, True, lanes=64)
        engs = [B200Engine(root, 3, True, lanes=64, runtime=CpuRuntime(), mode=m, adder_prefix=True) for m in ("levels", "serial")]
        for step in range(8):
            ora.step()
            for e in engs:
                e.step()
                for p in probes:
                    assert e.get_pin_state(p) == ora.get_pin_state(p), (seed, step, p, e.mode)
