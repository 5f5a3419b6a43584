"""not-gpu: serial programs (mcircuit_b200/serial.py) executed by the numpy test double, which
follows the stream kernel's execution model (a producer staging ring rows up to GB_MAX blocks ahead
and seeing only fenced stores; consumers with ACC / GREG / STASH / RING / LATE operands; CTA-wide
guard votes), checked against the oracle and the golden vectors."""
import numpy as np
import pytest

import mcircuit_b200.core.descriptors as MD
from _cpu_runtime import CpuRuntime
from _helpers import (build, compare_engine_to_oracle, golden, lane_trick_planes, run_register_case,
                      run_trace_case)
from mcircuit_b200 import circuits, compiler, serial
from mcircuit_b200.engine import B200Engine

HEAVY = ("c4_gate_level_lfsr32",)          # 97 probes x 400 steps: in full on the GPU and on the oracles, shortened on the numpy test double
ALL_TRACES = golden("traces.json")
TRACES = [c for c in ALL_TRACES if c["name"] not in HEAVY]


def serial_factory(keep="all", map_pins=True, **kw):
    def make(root, burst_size, lanes=64):
        return B200Engine(root, burst_size, map_pins, lanes=lanes, runtime=CpuRuntime(), mode="serial",
                          keep=keep, **kw)
    return make


@pytest.mark.parametrize("case", TRACES, ids=[c["name"] for c in TRACES])
@pytest.mark.parametrize("keep,map_pins,share", [("all", True, False), ("ports", True, True), ("all", False, False)])
def test_serial_traces(case, keep, map_pins, share):
    run_trace_case(case, serial_factory(keep, map_pins, share_edge_detectors=share),
                   skip_unobservable=(keep == "ports"))


@pytest.mark.parametrize("mode", ["serial", "levels"])
def test_heavy_traces_shortened(mode):
    """C4's shape at its own LFSR width out of gate-level flip-flops (unmodified reference JIT): first 48 steps here."""
    for case in ALL_TRACES:
        if case["name"] in HEAVY:
            short = dict(case, values=case["values"][:48], after_bursts=[])
            run_trace_case(short, lambda root, bs: B200Engine(root, bs, True, runtime=CpuRuntime(), mode=mode, keep="ports",
                                                              observe=[p for p in case["probes"]]), skip_unobservable=True)


@pytest.mark.parametrize("case", golden("register_intent.json"), ids=lambda c: c["name"])
def test_serial_register_intent(case):
    run_register_case(case, serial_factory())


@pytest.mark.parametrize("case", golden("adder.json"), ids=lambda c: c["name"])
def test_serial_adder(case):
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


@pytest.mark.parametrize("case", golden("lane_trick.json")[:5], ids=lambda c: c["name"])
def test_serial_lane_trick(case):
    ins, outs = lane_trick_planes(case)
    nw = case["n_words"]
    eng = B200Engine(build(case), 1, True, lanes=64 * nw * 3, runtime=CpuRuntime(), mode="serial", keep="ports")
    for p, words in ins.items():
        eng.set_pin_planes(p, np.tile(words, 3)[None, :])
    eng.step()
    for p, words in outs.items():
        assert np.array_equal(eng.get_pin_planes(p)[0], np.tile(words, 3)), p


@pytest.mark.parametrize("seed", range(100, 140))
def test_serial_vs_oracle_all_lanes(seed):
    prims = ("Gate", "Not", "Constant", "Clock", "Counter", "Counter", "Adder", "Register", "Register")
    root, probes = circuits.random_circuit(MD, seed, n_nodes=26, primitives=prims)
    d = compiler.flatten(root)
    lane_inputs = [(p, d.pin_width[d.find_pin(p)]) for p in ("/in0/pin", "/in1/pin")]
    compare_engine_to_oracle(lambda M: circuits.random_circuit(M, seed, n_nodes=26, primitives=prims),
                             lambda r, bs, lanes: serial_factory(share_edge_detectors=bool(seed % 2))(r, bs, lanes),
                             lanes=128, steps=8, probes=probes, lane_inputs=lane_inputs, seed=seed)


@pytest.mark.parametrize("keep", ["all", "ports"])
@pytest.mark.parametrize("share", [False, True])
def test_serial_lfsr_pipeline_vs_oracle(keep, share):
    """C4's shape: with shared edge detectors and keep='ports' nearly the whole step is one
    guarded region, skipped on every non-rising step - results must not change."""
    args = dict(n_lfsr=5, lfsr_bits=8, stages=40, taps=(7, 5, 1, 0))
    probes = [f"/tap{n}/pin" for n in range(5)] + ["/pipe_out/pin", "/clk/out"]
    if keep == "all":
        probes += [f"/l1_q{b}/out" for b in range(8)] + [f"/p{k}/out" for k in (0, 7, 39)] + ["/px3/out"]
    pokes = [(f"/l{n}_q{b}/out", (n + b) & 1) for n in range(5) for b in range(8)]
    eng, _ = compare_engine_to_oracle(lambda M: circuits.lfsr_pipeline(M, **args),
                                      lambda r, bs, lanes: serial_factory(keep, share_edge_detectors=share)(r, bs, lanes),
                                      lanes=64 * 3, steps=24, probes=probes, pokes=pokes)
    info = eng.program.serial_info
    if share and keep == "ports":
        assert info["regions"] == 1 and info["guarded_ops"] >= info["ops"] - 6
        assert eng.regions_skipped() == 12                   # every other step (the clock falls)
    assert info["acc_operands"] > 0 and info["stores_elided"] > 0


def gated_clock_chain(M):
    top = M.Composite()
    top.add_child("en", M.ExposedPin(M.ExposedPin.IN, 1))
    top.add_child("clk", M.Clock())
    top.add_child("gclk", M.Gate(M.Gate.AND, 1, 2))
    top.connect("clk", "out", "gclk", "in0")
    top.connect("en", "", "gclk", "in1")
    top.add_child("d", M.ExposedPin(M.ExposedPin.IN, 1))
    n = 48
    for i in range(n):
        top.add_child(f"r{i}", M.Register(1))
        top.add_child(f"x{i}", M.Gate(M.Gate.XOR, 1, 2))
    top.add_child("q", M.ExposedPin(M.ExposedPin.OUT, 1))
    for i in range(n):
        top.connect("gclk", "out", f"r{i}", "clock")
        top.connect("d" if i == 0 else f"r{i - 1}", "" if i == 0 else "out", f"x{i}", "in0")
        top.connect("d", "", f"x{i}", "in1")
        top.connect(f"x{i}", "out", f"r{i}", "data")
    top.connect(f"r{n - 1}", "out", "q", "")
    return top


def run_gated_clock_case(make_engine, lanes, idle_lanes, expect_skips):
    """The enable differs between lanes (a per-lane gated clock): CTAs whose lanes are all idle
    skip, the others execute - every lane must still match the oracle."""
    from oracle import restate
    en = np.zeros(lanes, dtype=np.uint64)
    en[idle_lanes:] = 1
    rng = np.random.default_rng(5)
    root = gated_clock_chain(MD)
    ora = restate.RestatedJIT(root, 1, True, lanes=lanes)
    eng = make_engine(root, lanes)
    assert eng.program.serial_info["regions"] >= 1
    for sim in (ora, eng):
        sim.set_pin_state("/en/pin", en)
    for step in range(12):
        d = rng.integers(0, 2, size=lanes, dtype=np.uint64)
        for sim in (ora, eng):
            sim.set_pin_state("/d/pin", d)
            sim.step()
        assert np.array_equal(eng.get_pin_state("/q/pin", lanes=True), ora.get_pin_state("/q/pin", lanes=True)), step
    assert eng.regions_skipped() == expect_skips


@pytest.mark.parametrize("model", ["early", "late"])
def test_serial_guard_is_per_cta_and_exact_with_lane_dependent_clocks(model):
    rt = CpuRuntime()
    rt.serial_cta_cols = 32                              # three "CTAs" of 32 lane words in the test double
    rt.serial_fetch_model = model
    # CTA 0 fully idle (never runs the region), CTA 1 mixed, CTA 2 fully active
    run_gated_clock_case(lambda root, lanes: B200Engine(root, 1, True, lanes=lanes, runtime=rt, mode="serial",
                                                        keep="ports", share_edge_detectors=True),
                         64 * 70, 64 * 40, 12)


def check_stream_invariants(prog):
    """The rules of serial.py's module docstring, re-derived by brute force from the encoded stream."""
    blocks = list(serial.decode(prog.records))
    nb = len(blocks)
    n_slots = prog.n_persist + prog.n_scratch
    stores = []
    for blk in blocks:
        st = set()
        for seg in blk["segments"]:
            f, n, r0 = seg["first_rec"], seg["count"], seg["first_row"]
            if seg["kind"] == serial.SEG_MUX:
                st.update(blk["rows"][r0:r0 + n])
            elif seg["kind"] == serial.SEG_XM:
                st.update(blk["rows"][r0:r0 + n] if seg["q_by_record"] else blk["rows"][r0 + n:r0 + 2 * n])
            elif seg["kind"] == serial.SEG_GENERIC:
                st.update(r["out"] for r in blk["records"][f:f + n] if r["store"])
        stores.append(st)
    for b, blk in enumerate(blocks):
        assert blk["n_rows"] <= serial.RING_ROWS and len(blk["runs"]) <= serial.MAX_RUNS
        assert len(blk["segments"]) <= serial.MAX_SEG and len(blk["records"]) <= serial.BLOCK
        assert blk["fence"] == (b % serial.FENCE_PERIOD == 0)
        covered = [run["slot"] + k for run in blk["runs"] for k in range(run["rows"])]
        assert covered == list(blk["rows"]) and all(s < n_slots for s in covered)
        for run in blk["runs"]:                          # a TMA box lies in one plane
            assert (run["slot"] < prog.n_persist) == (run["slot"] + run["rows"] - 1 < prog.n_persist)
        stored = set()
        seen_recs = 0
        for k, seg in enumerate(blk["segments"]):
            f, n, r0 = seg["first_rec"], seg["count"], seg["first_row"]
            assert f == seen_recs
            if seg["kind"] == serial.SEG_GUARD:
                assert k == len(blk["segments"]) - 1 and blk["guard"] and b + 1 + blk["span"] <= nb
                r = blk["records"][f]
                if r["kinds"][0] == serial.K_RING:
                    assert blk["rows"][r["fields"][0]] not in stored
                seen_recs += 1
                continue
            if seg["kind"] == serial.SEG_MUX:
                r = blk["records"][f]
                if r["kinds"][1] == serial.K_RING:
                    assert blk["rows"][r["fields"][1]] not in stored
                for u in range(n):
                    rec = blk["records"][f + u]
                    assert rec["op"] == 6 and rec["store"] and rec["out"] == blk["rows"][r0 + u]
                    assert rec["stash_out"] == (seg["stash_out"] if u == n - 1 else -1)
                    assert blk["rows"][r0 + u] not in stored, "ring row stored earlier in its block"
                    stored.add(rec["out"])
                seen_recs += n
            elif seg["kind"] == serial.SEG_XM:
                for u in range(n):
                    old_row = r0 + u if seg["q_by_record"] else r0 + n + u
                    x = blk["records"][f + 2 * u]
                    if seg["q_by_record"]:
                        assert x["kinds"][1] in (serial.K_STASH, serial.K_LATE) and x["kinds"][0] == serial.K_ACC
                    else:
                        assert blk["rows"][r0 + u] not in stored
                    assert blk["rows"][old_row] not in stored
                    assert blk["records"][f + 2 * u + 1]["out"] == blk["rows"][old_row]
                    assert blk["records"][f + 2 * u + 1]["stash_out"] == (seg["stash_out"] if u == n - 1 else -1)
                    stored.add(blk["rows"][old_row])
                seen_recs += 2 * n
            else:
                for rec in blk["records"][f:f + n]:
                    for kind, fl in zip(rec["kinds"], rec["fields"]):
                        if kind == serial.K_RING:
                            assert fl < blk["n_rows"] and blk["rows"][fl] not in stored, "ring row stored earlier in its block"
                        if kind == serial.K_LATE:
                            assert fl < n_slots
                        if kind == serial.K_STASH:
                            assert fl < serial.STASH_N
                    if rec["store"]:
                        assert rec["out"] < n_slots
                        stored.add(rec["out"])
                seen_recs += n
        assert seen_recs == len(blk["records"])
        assert not blk["guard"] or blk["segments"][-1]["kind"] == serial.SEG_GUARD
        # no store to a staged row between the point that publishes stores to the producer and this block
        k = b - 1
        scanned = 0
        while scanned < 4 * nb + serial.GB_MAX + serial.FENCE_PERIOD:
            kb = k % nb
            if blocks[kb]["guard"]:
                break                                    # a guard publishes its own block's stores too
            for s in blk["rows"]:
                assert s not in stores[kb], ("block %d stages slot %d that block %d stores" % (b, s, kb))
            if blocks[kb]["fence"] and (b - k) >= serial.GB_MAX:
                break                                    # the fence at its start published everything before kb
            k -= 1
            scanned += 1


def test_serial_stream_invariants():
    for share in (False, True):
        check_stream_invariants(serial.compile_circuit_serial(
            circuits.lfsr_pipeline(MD, 4, 8, 30, taps=(7, 5, 1, 0)), keep="ports", share_edge_detectors=share))
    check_stream_invariants(serial.compile_circuit_serial(circuits.lfsr_pipeline(MD, 40, 16, 300), keep="ports",
                                                          share_edge_detectors=True))
    for seed in range(300, 310):
        root, _ = circuits.random_circuit(MD, seed, n_nodes=40)
        check_stream_invariants(serial.compile_circuit_serial(root))
    check_stream_invariants(serial.compile_circuit_serial(circuits.ripple_adder(MD, 32), keep="ports"))
    check_stream_invariants(serial.compile_circuit_serial(circuits.counter_circuit(MD)))


@pytest.mark.parametrize("model", ["early", "late"])
@pytest.mark.parametrize("options", [{}, {"stash": False}, {"ring": False}, {"patterns": False}, {"stash": False, "ring": False}])
def test_serial_many_steps_in_one_launch(model, options):
    """The producer runs ahead across the end of a step: run(k) in ONE call against the oracle stepping."""
    from oracle import restate
    rt = CpuRuntime()
    rt.serial_fetch_model = model
    cases = [(circuits.lfsr_pipeline(MD, 5, 8, 40, taps=(7, 5, 1, 0)), True), (circuits.counter_circuit(MD), False),
             (circuits.raw_counter(MD, 3, 2), False), (circuits.gate_level_lfsr(MD), False)]
    for seed in (400, 401, 402, 403):
        cases.append((circuits.random_circuit(MD, seed, n_nodes=30, primitives=(
            "Gate", "Not", "Constant", "Clock", "Counter", "Counter", "Adder", "Register", "Register"))[0], bool(seed % 2)))
    for root, share in cases:
        eng = B200Engine(root, 3, True, lanes=128, runtime=rt, mode="serial", keep="all", share_edge_detectors=share,
                         serial_options=options)
        ora = restate.RestatedJIT(root, 3, True, lanes=128)
        d = compiler.flatten(root)
        pins = [p[0] for p in d.iter_pins()]
        rng = np.random.default_rng(11)
        for p in pins:                                   # an all-zero LFSR stays zero and proves nothing
            if p.endswith("/out") and "_q" in p:
                v = rng.integers(0, 2, size=128, dtype=np.uint64)
                eng.set_pin_state(p, v)
                ora.set_pin_state(p, v)
        compared = 0
        for chunk in (1, 2, 5, 17, 64):
            eng.run(chunk)
            for _ in range(chunk):
                ora.step()
            for p in pins:
                assert np.array_equal(eng.get_pin_state(p, lanes=True), ora.get_pin_state(p, lanes=True)), (chunk, p, options)
                compared += 1
        assert compared >= 5 * 3


def test_compiling_one_design_twice_with_shared_edge_detectors_is_stable():
    """lower(share_edge_detectors=True) used to rewrite design.net_sig in place, so a
    second compile of the same Design put the leader's prevclock into recyclable scratch."""
    design = compiler.flatten(circuits.lfsr_pipeline(MD, 3, 8, 10, taps=(7, 5, 1, 0)))
    before = design.net_sig.copy()
    a = compiler.compile_design(design, keep="ports", share_edge_detectors=True)
    b = compiler.compile_design(design, keep="ports", share_edge_detectors=True)
    c = compiler.compile_design(design, keep="ports", share_edge_detectors=False)
    assert np.array_equal(design.net_sig, before)
    assert (a.n_persist, a.n_scratch, a.n_gates) == (b.n_persist, b.n_scratch, b.n_gates)
    assert np.array_equal(a.records, b.records) and np.array_equal(a.sig_slot, b.sig_slot)
    assert np.array_equal(a.sig_kept, b.sig_kept)
    assert c.n_gates > a.n_gates
    lead = a.pin_slots("/l0_q0/prevclock")[0]                  # whichever pin leads, members view its slot
    assert all(a.pin_slots(f"/l{n}_q{k}/prevclock")[0] == lead for n in range(3) for k in range(8))
    assert lead < a.n_persist
