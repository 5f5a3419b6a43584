"""not-gpu: the SURVEY 8f "next" items - headless Schematic.reconstruct composites, the
netlist file format / compiled-program cache and waveform capture - on the numpy test double."""
import json

import numpy as np
import pytest

import mcircuit_b200.core.descriptors as MD
from _cpu_runtime import CpuRuntime
from mcircuit_b200 import circuits, compiler, netlist_io, schematic, waveform
from mcircuit_b200.engine import B200Engine
from oracle import restate
from _helpers import golden, run_schematic_case


def _gui_full_adder(M=MD):
    """What the GUI would hold for a 1-bit full adder drawn with LIBRARY elements (app.py:575-589)."""
    return circuits.gui_full_adder(M)


def test_schematic_reconstruct_shape_and_result():
    elements, nets = _gui_full_adder()
    s = schematic.reconstruct(elements, nets)
    assert list(s.graph.nodes) == [n for n, _ in elements]          # children in element order
    assert ("input_1", "pin", "xor_4", "in0") in s.connections and ("or_8", "out", "output_10", "pin") in s.connections
    assert len(s.connections) == 12
    # sources outer, dests inner: first edges come from input0 in dest order
    assert list(s.graph.edges)[:2] == [("xor_4", "input_1"), ("xor_4", "input_2")] or \
        list(s.graph.predecessors("input_1")) == ["xor_4", "and_7"]
    # runs through the engine exactly like app.py:732 does: JIT(s, 500, True)
    eng = B200Engine(s, 500, True, runtime=CpuRuntime())
    ora = restate.RestatedJIT(s, 500, True)
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                for sim in (eng, ora):
                    sim.set_pin_state("/input_1/pin", a)
                    sim.set_pin_state("/input_2/pin", b)
                    sim.set_pin_state("/input_3/pin", c)
                    sim.step()
                assert eng.get_pin_state("/output_9/pin") == ora.get_pin_state("/output_9/pin") == (a ^ b ^ c)
                assert eng.get_pin_state("/output_10/pin") == ora.get_pin_state("/output_10/pin") == (a + b + c) // 2


def test_schematic_matches_live_reference_connect_order(reference):
    D, S = reference
    el_ref, nets = _gui_full_adder(D)
    s_ref = schematic.reconstruct(el_ref, nets, M=D)
    s_own = schematic.reconstruct(*_gui_full_adder(MD))
    emits_ref = [v[1] for k, v in S.iter_simulation_topology(s_ref) if k == "emit"]
    d = compiler.flatten(s_own)
    assert emits_ref == [d.leaf_path[a] for k, a in d.schedule if k == 0]


@pytest.mark.parametrize("seed", range(12))
def test_netlist_json_roundtrip_preserves_order_and_results(tmp_path, seed):
    root, probes = circuits.random_circuit(MD, seed, n_nodes=20,
                                           primitives=("Gate", "Not", "Constant", "Clock", "Counter", "Adder", "Register"))
    path = tmp_path / "n.json"
    netlist_io.save(root, path)
    back = netlist_io.load(path)
    a, b = compiler.flatten(root), compiler.flatten(back)
    assert list(a.iter_pins()) == list(b.iter_pins())
    assert a.schedule == b.schedule or [x for x in a.schedule if x[0] == 0] == [x for x in b.schedule if x[0] == 0]
    assert [a.leaf_path[x] for k, x in a.schedule if k == 0] == [b.leaf_path[x] for k, x in b.schedule if k == 0]
    assert np.array_equal(a.pin_src, b.pin_src)
    pa, pb = compiler.compile_design(a), compiler.compile_design(b)
    assert np.array_equal(pa.records, pb.records) and np.array_equal(pa.level_off, pb.level_off)
    json.loads(path.read_text())["mcircuit_b200_netlist"] == 1


def test_netlist_json_shares_instanced_descriptors(tmp_path):
    root = circuits.raw_counter(MD, 2, 1)
    d = netlist_io.to_dict(root)
    assert sum(1 for x in d["defs"] if x["type"] == "Composite") == 6      # adder clock sr d latch main
    back = netlist_io.from_dict(d)
    assert back.get_child("b1") is back.get_child("b2")                    # still one object, two names


def test_program_cache_runs_without_design(tmp_path):
    root = circuits.random_dag(MD, n_gates=800, depth=16, n_inputs=32, seed=2)
    prog = compiler.compile_circuit(root, keep="ports")
    path = tmp_path / "prog.npz"
    netlist_io.save_program(prog, path)
    cached = netlist_io.load_program(path)
    a = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), keep="ports", mode="levels", program=prog)
    b = B200Engine(None, 1, True, lanes=128, runtime=CpuRuntime(), mode="levels", program=cached)
    for e in (a, b):
        e.randomize_inputs()
        e.step()
    assert b.input_pins() == a.input_pins() and b.output_pins() == a.output_pins()
    for p in a.output_pins():
        assert np.array_equal(a.get_pin_planes(p), b.get_pin_planes(p))
    with pytest.raises(KeyError):
        b.get_pin_state("/g3_7/out")                                        # not retained -> not in the cache


def test_capture_and_vcd():
    eng = B200Engine(circuits.counter_circuit(MD), 1, True, lanes=128, runtime=CpuRuntime())
    pins = ["/clk/out", "/r/out"]
    samples = eng.capture(pins, 8)
    assert samples.shape == (8, 17, 2)
    clk, r = waveform.lane_values(samples, eng.pin_widths(pins), lane=70)
    assert clk == [1, 0, 1, 0, 1, 0, 1, 0] and r == [1, 1, 0, 0, 1, 1, 0, 0]
    vcd = waveform.to_vcd(pins, eng.pin_widths(pins), samples)
    assert "$var wire 1 ! clk.out $end" in vcd and '$var wire 16 " r.out $end' in vcd
    assert '#0\n1!\nb1 "' in vcd and '#2\n1!\nb0 "' in vcd
    assert eng.steps_done == 8


def test_netlist_file_loads_into_the_reference_too(reference, tmp_path):
    """A netlist saved from our descriptors rebuilds with the REFERENCE's descriptor module and
    runs on its JIT with the same results as our engine on the original."""
    from oracle import refshim
    D, _ = reference
    root, probes = circuits.random_circuit(MD, 77, n_nodes=18)
    path = tmp_path / "n.json"
    netlist_io.save(root, path)
    ref_root = netlist_io.load(path, M=D)
    jit = refshim.build_jit(ref_root, 3, True)
    eng = B200Engine(root, 3, True, runtime=CpuRuntime())
    for step in range(8):
        jit.step()
        eng.step()
        for p in probes:
            assert jit.get_pin_state(p) == eng.get_pin_state(p), (step, p)


def test_checkpoint_resume():
    """state_dict / load_state_dict: a run resumed from a checkpoint continues bit-exactly."""
    root = circuits.lfsr_pipeline(MD, 3, 8, 6, taps=(7, 5, 1, 0))
    a = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), mode="resident", keep="ports")
    for n in range(3):
        a.set_pin_state(f"/l{n}_q0/out", 1)
    a.run(11)
    snap = a.state_dict()
    a.run(20)
    b = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), mode="levels", keep="ports")
    b.load_state_dict(snap)
    assert b.steps_done == 11
    b.run(20)
    for p in a.output_pins():
        assert np.array_equal(a.get_pin_planes(p), b.get_pin_planes(p)), p
    other = B200Engine(circuits.counter_circuit(MD), 1, True, lanes=128, runtime=CpuRuntime())
    with pytest.raises(ValueError):
        other.load_state_dict(snap)


@pytest.mark.parametrize("case", golden("schematic.json"), ids=lambda c: c["name"])
def test_schematic_golden_from_reference_jit(case, tmp_path):
    """Sheets rebuilt the way Schematic.reconstruct does, against values the reference JIT produced;
    the oracle, the engine (test double) and the engine on the JSON-saved-and-reloaded sheet."""
    run_schematic_case(case, lambda root, bs: restate.RestatedJIT(root, bs, True))
    run_schematic_case(case, lambda root, bs: B200Engine(root, bs, True, runtime=CpuRuntime()))
    path = tmp_path / "sheet.json"
    netlist_io.save(circuits.gui_circuit(MD, case["sheet"]), path)
    run_schematic_case(case, lambda root, bs: B200Engine(root, bs, True, runtime=CpuRuntime()), root=netlist_io.load(path))


@pytest.mark.parametrize("modes", [("levels", "serial"), ("serial", "resident"), ("resident", "serial"), ("serial", "levels")])
@pytest.mark.parametrize("back_edges", ["auto", "latch"])
def test_checkpoint_resume_across_execution_modes(modes, back_edges):
    """A checkpoint is keyed by signal: written under one slot layout, resumed under another."""
    root = circuits.lfsr_pipeline(MD, 3, 8, 6, taps=(7, 5, 1, 0))
    a = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), mode=modes[0], keep="ports", back_edges=back_edges)
    ora = restate.RestatedJIT(root, 1, True, lanes=128)
    for n in range(3):
        a.set_pin_state(f"/l{n}_q0/out", 1)
        ora.set_pin_state(f"/l{n}_q0/out", 1)
    a.run(11)
    snap = a.state_dict()
    b = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), mode=modes[1], keep="ports", back_edges=back_edges)
    b.load_state_dict(snap)
    b.run(20)
    for _ in range(31):
        ora.step()
    for p in b.output_pins():
        assert np.array_equal(b.get_pin_state(p, lanes=True), ora.get_pin_state(p, lanes=True)), p
    # ... and across share_edge_detectors: checkpoints are keyed by net bit, members' prevclock nets resolve to the leader
    for share_a, share_b in ((False, True), (True, False)):
        a = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), mode=modes[0], keep="ports", share_edge_detectors=share_a)
        ora = restate.RestatedJIT(root, 1, True, lanes=128)
        for n in range(3):
            a.set_pin_state(f"/l{n}_q0/out", 1)
            ora.set_pin_state(f"/l{n}_q0/out", 1)
        a.run(7)
        b = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), mode=modes[1], keep="ports", share_edge_detectors=share_b)
        b.load_state_dict(a.state_dict())
        b.run(12)
        for _ in range(19):
            ora.step()
        for p in b.output_pins() + ["/l1_q3/prevclock", "/p2/prevclock"]:
            assert np.array_equal(b.get_pin_state(p, lanes=True), ora.get_pin_state(p, lanes=True)), (p, share_a, share_b)


@pytest.mark.parametrize("mode", ["auto", "levels", "serial", "resident"])
def test_poking_one_members_prevclock_unshares_the_engine(mode):
    """Edge detectors are shared by default; the reference has one prevclock global per primitive
    (simulator.py:100-127), so a poke of ONE member's prevclock rebuilds the engine unshared, state intact."""
    root = circuits.lfsr_pipeline(MD, 3, 8, 6, taps=(7, 5, 1, 0))
    eng = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), mode=mode)
    ora = restate.RestatedJIT(root, 1, True, lanes=128)
    assert eng.program.shared_prev
    n_shared = eng.program.n_gates
    rng = np.random.default_rng(3)
    for n in range(3):
        v = rng.integers(0, 2, size=128, dtype=np.uint64)
        eng.set_pin_state(f"/l{n}_q0/out", v)
        ora.set_pin_state(f"/l{n}_q0/out", v)
    eng.run(5)
    for _ in range(5):
        ora.step()
    assert eng.program.shared_prev                       # ordinary pokes and runs keep sharing
    lanes_v = rng.integers(0, 2, size=128, dtype=np.uint64)
    for sim in (eng, ora):
        sim.set_pin_state("/l1_q4/prevclock", lanes_v)     # one register of the group, per lane
        sim.set_pin_state("/p3/prevclock", 1)
    assert not eng.program.shared_prev and eng.program.n_gates > n_shared and eng.steps_done == 5
    for step in range(9):
        eng.step()
        ora.step()
        for p in eng.output_pins() + ["/l1_q4/out", "/l1_q4/prevclock", "/l1_q5/prevclock", "/p3/out", "/p3/prevclock"]:
            assert np.array_equal(eng.get_pin_state(p, lanes=True), ora.get_pin_state(p, lanes=True)), (step, p)
    with pytest.raises(KeyError):
        eng.set_pin_state("/nope/prevclock", 1)


def test_auto_mode_policy():
    """mode='auto' (measured in profiles/r02_modes_compare.jsonl): register-dominated sequential designs run on the
    stream kernel, small gate-level designs on the resident kernel, wide netlists level by level."""
    def mode_of(root, lanes):
        return B200Engine(root, 1, True, lanes=lanes, runtime=CpuRuntime(), keep="ports").mode
    assert mode_of(circuits.lfsr_pipeline(MD, 8, 16, 60, taps=(15, 13, 12, 10)), 64 * 16) == "serial"
    assert mode_of(circuits.gate_level_lfsr(MD, n_bits=32, taps=(31, 21, 1, 0), stages=64), 64 * 16) == "serial"   # forwarded
    assert mode_of(circuits.gate_level_lfsr(MD), 64 * 16) == "resident"        # too short to stage its state: late loads
    assert mode_of(circuits.counter_circuit(MD), 64) == "resident"
    assert mode_of(circuits.array_multiplier(MD, 16), 64 * 16) == "resident"   # scattered operands: a box per operand
    assert mode_of(circuits.random_dag(MD, n_gates=6000, depth=12, n_inputs=64, seed=5), 64 * 64) == "levels"


def test_prevclock_poke_with_toggle_counting_is_refused_cleanly():
    root = circuits.lfsr_pipeline(MD, 3, 8, 6, taps=(7, 5, 1, 0))
    eng = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), mode="serial")
    eng.enable_toggle_counts(["/tap0/pin"])
    eng.run(3)
    with pytest.raises(compiler.CircuitError):
        eng.set_pin_state("/l1_q4/prevclock", 1)
    assert eng.program.shared_prev and eng.steps_done == 3          # nothing was rebuilt
    eng.run(2)                                                       # and the engine still works


def test_program_cache_of_a_stream_program(tmp_path):
    """A cached serial (K3) program keeps its kind: it runs in mode 'serial' and gives the same results."""
    from mcircuit_b200 import serial
    root = circuits.lfsr_pipeline(MD, 4, 8, 20, taps=(7, 5, 1))
    prog = serial.compile_circuit_serial(root, keep="ports", share_edge_detectors=True)
    path = tmp_path / "stream.npz"
    netlist_io.save_program(prog, path)
    cached = netlist_io.load_program(path)
    assert cached.serial and cached.records.shape == prog.records.shape
    a = B200Engine(root, 1, True, lanes=128, runtime=CpuRuntime(), keep="ports", program=prog)
    b = B200Engine(None, 1, True, lanes=128, runtime=CpuRuntime(), program=cached)
    assert a.mode == b.mode == "serial"
    with pytest.raises(ValueError):
        B200Engine(None, 1, True, lanes=128, runtime=CpuRuntime(), program=cached, mode="levels")
    for e in (a, b):
        e.randomize_inputs(["/l%d_q0/out" % n for n in range(4)])
        e.run(25)
    for p in a.output_pins():
        assert np.array_equal(a.get_pin_planes(p), b.get_pin_planes(p)), p
    assert int(np.unpackbits(a.get_pin_planes("/pipe_out/pin").view(np.uint8)).sum()) > 0
