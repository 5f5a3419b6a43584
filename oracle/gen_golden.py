"""Generate tests/golden/*.json by RUNNING THE UNMODIFIED REFERENCE (its LLVM JIT) in this
container.  TEST INFRASTRUCTURE - run by hand: ``python oracle/gen_golden.py``.

The reference is Python and cannot travel to the GPU box, so its outputs are committed as
small fixtures.  Every case records how the circuit is built (a builder of
``mcircuit_b200.circuits`` called with the reference's own descriptor module), the pokes,
and the values the reference's JIT produced; tests rebuild the same circuit with our
descriptors and compare the oracle (CPU) and the CUDA engine (GPU) against them.

"lane trick" cases: a circuit made only of ExposedPin/Constant/Not/Gate is
built with every descriptor at width=64 and fed packed 64-lane words; because gates are
bitwise (simulator.py:83-97) the JIT's output words are exactly the bit-sliced words the
engine must produce for the width=1 circuit on those 64 lanes.

``register_intent`` cases are NOT reference results: the reference's Register emitter does
not compile (simulator.py:115-127).  They come from the reference JIT with only that one
emitter replaced, at run time, by the evident intent, and are labelled as such.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from mcircuit_b200 import circuits  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
D, S = refshim.load()


def trace(root, probes, steps, pokes=(), burst_size=7, map_pins=True, bursts_after=0):
    jit = refshim.build_jit(root, burst_size, map_pins)
    for p, v in pokes:
        jit.set_pin_state(p, v)
    rows = []
    for _ in range(steps):
        jit.step()
        rows.append([jit.get_pin_state(p) for p in probes])
    after = []
    for _ in range(bursts_after):
        jit.burst()
        after.append([jit.get_pin_state(p) for p in probes])
    return rows, after


def case_trace(name, builder, args, probes, steps, pokes=(), burst_size=7, bursts_after=0,
               both_modes=True, note=""):
    out = {"name": name, "kind": "trace", "builder": builder, "args": args, "probes": probes,
           "steps": steps, "pokes": [list(p) for p in pokes], "burst_size": burst_size,
           "bursts_after": bursts_after, "note": note}
    made = getattr(circuits, builder)(D, **args)
    root = made[0] if isinstance(made, tuple) else made
    rows, after = trace(root, probes, steps, pokes, burst_size, True, bursts_after)
    out["values"] = rows
    out["after_bursts"] = after
    if both_modes:
        rows2, after2 = trace(root, probes, steps, pokes, burst_size, False, bursts_after)
        assert rows2 == rows and after2 == after, "map_pins modes disagree in the reference?!"
    return out


def adder_vectors(width):
    m = (1 << width) - 1
    h = 1 << (width - 1)
    vals = sorted({0, 1, m, h, h - 1 if h > 1 else 0, m - 1 if m > 1 else 0, (0x5555555555555555 & m), (0xAAAAAAAAAAAAAAAA & m)})
    rng = random.Random(width)
    vecs = [(a, b, c) for a in vals for b in vals for c in (0, 1)]
    for _ in range(24):
        vecs.append((rng.getrandbits(width), rng.getrandbits(width), rng.getrandbits(1)))
    if width <= 3:
        vecs = [(a, b, c) for a in range(m + 1) for b in range(m + 1) for c in (0, 1)]
    return vecs


def case_adder(width):
    top = D.Composite()
    top.add_child("add", D.Adder(width))
    jit = refshim.build_jit(top, 1, True)
    vecs, res = adder_vectors(width), []
    for a, b, c in vecs:
        jit.set_pin_state("/add/a", a)
        jit.set_pin_state("/add/b", b)
        jit.set_pin_state("/add/cin", c)
        jit.step()
        res.append([jit.get_pin_state("/add/sum"), jit.get_pin_state("/add/cout")])
    return {"name": "adder_w%d" % width, "kind": "adder", "width": width,
            "vectors": [list(v) for v in vecs], "results": res}


def case_lane_trick(name, builder, args, in_pins, out_pins, n_words, seed):
    root = getattr(circuits, builder)(D, width=64, **args)
    jit = refshim.build_jit(root, 1, True)
    rng = random.Random(seed)
    ins = {p: [] for p in in_pins}
    outs = {p: [] for p in out_pins}
    for _ in range(n_words):
        for p in in_pins:
            w = rng.getrandbits(64)
            ins[p].append(w)
            jit.set_pin_state(p, w)
        jit.step()
        for p in out_pins:
            outs[p].append(jit.get_pin_state(p))
    return {"name": name, "kind": "lane_trick", "builder": builder, "args": args,
            "n_words": n_words, "inputs": ins, "outputs": outs}


def patch_register_intent():
    """Replace ONLY the Register emitter of the loaded reference module by the evident intent
    of simulator.py:115-127: sign-extend the rising edge to a mask and load ``out``."""
    import llvmlite.ir as ll

    def intent(b, desc, path, get_global):
        prevclk = get_global(path, "prevclock")
        data = get_global(path, "data")
        out = get_global(path, "out")
        clk = get_global(path, "clock")
        itype = ll.IntType(desc.width)
        v3 = b.and_(b.not_(b.load(prevclk)), b.load(clk))
        mask = b.sext(v3, itype) if desc.width > 1 else v3
        b.store(b.or_(b.and_(mask, b.load(data)), b.and_(b.not_(mask), b.load(out))), out)
        b.store(b.load(clk), prevclk)

    S.TRANSLATOR[D.Register] = intent


def register_shift(M, width=1, n=4):
    top = M.Composite()
    top.add_child("clk", M.Clock())
    top.add_child("d", M.ExposedPin(M.ExposedPin.IN, width))
    for i in range(n):
        top.add_child(f"r{i}", M.Register(width))
    for i in range(n):
        top.connect("clk", "out", f"r{i}", "clock")
        top.connect("d" if i == 0 else f"r{i - 1}", "" if i == 0 else "out", f"r{i}", "data")
    return top


def schematic_cases():
    """GUI-shaped sheets (mcircuit_b200/circuits.py::gui_*) rebuilt with the REFERENCE's descriptor
    module the way Schematic.reconstruct (diagram.py:88-119) builds them, run on the reference JIT
    the way the GUI does (app.py:732: JIT(s, 500, True), then set_pin_state / step)."""
    out = []
    root = circuits.gui_circuit(D, "full_adder")
    jit = refshim.build_jit(root, 500, True)
    probes = ["/output_9/pin", "/output_10/pin", "/xor_4/out", "/nand_11/out"]
    script, rows = [], []
    for v in range(8):
        pokes = [["/input_1/pin", v & 1], ["/input_2/pin", (v >> 1) & 1], ["/input_3/pin", (v >> 2) & 1]]
        for p, x in pokes:
            jit.set_pin_state(p, x)
        jit.step()
        script.append(pokes)
        rows.append([jit.get_pin_state(p) for p in probes])
    out.append({"name": "gui_full_adder", "kind": "schematic", "sheet": "full_adder", "burst_size": 500,
                "probes": probes, "script": script, "values": rows})
    root = circuits.gui_circuit(D, "latch_and_ring")
    jit = refshim.build_jit(root, 500, True)
    probes = ["/output_9/pin", "/output_10/pin", "/output_11/pin", "/not_5/out", "/not_6/out", "/not_7/out"]
    rng = random.Random(5)
    script, rows = [], []
    for step in range(40):
        pokes = []
        if step % 5 == 0:
            r = rng.getrandbits(2)
            pokes = [["/input_1/pin", r & 1], ["/input_2/pin", (r >> 1) & 1 if (r & 1) == 0 else 0]]
        for p, x in pokes:
            jit.set_pin_state(p, x)
        jit.step()
        script.append(pokes)
        rows.append([jit.get_pin_state(p) for p in probes])
    jit.burst()
    after = [jit.get_pin_state(p) for p in probes]
    out.append({"name": "gui_latch_and_ring", "kind": "schematic", "sheet": "latch_and_ring", "burst_size": 500,
                "probes": probes, "script": script, "values": rows, "after_burst": after})
    with open(os.path.join(OUT, "schematic.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("schematic.json:", len(out), "cases")


def c4_gate_level_case():
    """C4's shape at C4's own LFSR width and taps (32 bits, x^32 + x^22 + x^2 + x + 1) plus a 64-stage pipeline, built from
    the gate-level master-slave flip-flops of examples/raw_counter.py, which the UNMODIFIED reference compiles and runs."""
    return case_trace("c4_gate_level_lfsr32", "gate_level_lfsr", {"n_bits": 32, "taps": [31, 21, 1, 0], "stages": 64},
                      [f"/q{i}/pin" for i in range(32)] + [f"/p{k}/pin" for k in range(64)] + ["/clk/out/pin"], 400,
                      note="32-bit LFSR with C4's taps + 64-stage pipeline out of gate-level flip-flops; unmodified reference JIT")


def main():
    os.makedirs(OUT, exist_ok=True)
    schematic_cases()
    if "--schematic-only" in sys.argv:
        return
    if "--c4-gate-level-only" in sys.argv:                 # add / refresh that one case, keep the rest of traces.json
        path = os.path.join(OUT, "traces.json")
        cases = [c for c in json.load(open(path)) if c["name"] != "c4_gate_level_lfsr32"]
        cases.append(c4_gate_level_case())
        with open(path, "w") as f:
            json.dump(cases, f, separators=(",", ":"))
        print("traces.json:", len(cases), "cases")
        return
    cases = []
    # --- the reference's three examples ---------------------------------
    cases.append(case_trace("examples_clock", "inverter_clock", {}, ["/not/out"], 10))
    cases.append(case_trace("examples_counter", "counter_circuit", {},
                            ["/clk/out", "/r/out", "/r/prevclock"], 8, burst_size=501, bursts_after=2))
    cases.append(case_trace("counter_w1", "counter_circuit", {"width": 1},
                            ["/clk/out", "/r/out", "/r/prevclock"], 8))
    for seed in (0, 1, 2, 3):
        cases.append(case_trace("examples_raw_counter_seed%d" % seed, "raw_counter",
                                {"bits": 2, "seed": seed}, ["/b1/q/pin", "/b2/q/pin"], 24,
                                pokes=[("/a1/cin/pin", 1)]))
    cases.append(case_trace("raw_counter_4bit", "raw_counter", {"bits": 4, "seed": 7},
                            ["/b1/q/pin", "/b2/q/pin", "/b3/q/pin", "/b4/q/pin"], 80,
                            pokes=[("/a1/cin/pin", 1)]))
    # --- C4 as the UNMODIFIED reference can run it: gate-level flip-flops ---
    cases.append(case_trace("c4_gate_level_lfsr", "gate_level_lfsr", {"n_bits": 8, "taps": [7, 5, 4, 3], "stages": 4},
                            [f"/q{i}/pin" for i in range(8)] + [f"/p{k}/pin" for k in range(4)] + ["/clk/out/pin"], 200,
                            note="LFSR + pipeline out of the master-slave flip-flops of examples/raw_counter.py"))
    cases.append(c4_gate_level_case())
    # --- random circuits: every primitive the reference can compile, feedback, hierarchy -------
    for seed in range(40):
        made, probes = circuits.random_circuit(D, seed, n_nodes=18)
        cases.append(case_trace("fuzz_seed%d" % seed, "random_circuit", {"seed": seed, "n_nodes": 18},
                                probes, 10))
    with open(os.path.join(OUT, "traces.json"), "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    print("traces.json:", len(cases), "cases")

    # --- Adder semantics -------------------------------------------------------------
    adders = [case_adder(w) for w in (1, 2, 3, 4, 8, 16, 32, 63, 64)]
    with open(os.path.join(OUT, "adder.json"), "w") as f:
        json.dump(adders, f, separators=(",", ":"))
    print("adder.json:", sum(len(a["vectors"]) for a in adders), "vectors")

    # --- lane trick: C2 adder and small C5-style DAGs ----------------------------------------------
    lane = []
    nb = 32
    lane.append(case_lane_trick(
        "c2_ripple_adder32", "ripple_adder", {"nbits": nb},
        ["/cin/pin"] + [f"/a{i}/pin" for i in range(nb)] + [f"/b{i}/pin" for i in range(nb)],
        [f"/s{i}/pin" for i in range(nb)] + ["/cout/pin"], 16, 1))
    lane.append(case_lane_trick(
        "c3_array_multiplier8", "array_multiplier", {"n": 8},
        [f"/a{i}/pin" for i in range(8)] + [f"/b{i}/pin" for i in range(8)],
        [f"/p{i}/pin" for i in range(16)], 8, 2))
    for gates, depth, nin, seed in ((400, 10, 32, 11), (1500, 30, 64, 12)):
        lane.append(case_lane_trick(
            "c5_random_dag_%d" % gates, "random_dag",
            {"n_gates": gates, "depth": depth, "n_inputs": nin, "seed": seed},
            [f"/i{k}/pin" for k in range(nin)], [f"/o{k}/pin" for k in range(gates // depth)], 4, seed))
    if "--full-c3" in sys.argv:
        # the whole C3 netlist on the reference JIT: ~2.5 minutes of JIT build
        lane.append(case_lane_trick(
            "c3_array_multiplier64", "array_multiplier", {"n": 64},
            [f"/a{i}/pin" for i in range(64)] + [f"/b{i}/pin" for i in range(64)],
            [f"/p{i}/pin" for i in range(128)], 2, 64))
        # the C5 generator at 10^4 gates: ~2 minutes of JIT build, the practical limit
        args = {"n_gates": 10000, "depth": 50, "n_inputs": 256, "seed": 13}
        lane.append(case_lane_trick(
            "c5_random_dag_10000", "random_dag", args,
            [f"/i{k}/pin" for k in range(256)], [f"/o{k}/pin" for k in range(200)], 2, 5))
    else:
        # keep the committed slow cases when they are not being regenerated
        try:
            with open(os.path.join(OUT, "lane_trick.json")) as f:
                lane += [c for c in json.load(f) if c["name"] in ("c3_array_multiplier64", "c5_random_dag_10000")]
        except OSError:
            pass
    with open(os.path.join(OUT, "lane_trick.json"), "w") as f:
        json.dump(lane, f, separators=(",", ":"))
    print("lane_trick.json:", len(lane), "cases")

    # --- Register: intent, not a reference result ----------------------------------------------------
    patch_register_intent()
    regs = []
    for width, n in ((1, 4), (8, 3)):
        root = register_shift(D, width, n)
        probes = [f"/r{i}/out" for i in range(n)] + [f"/r{i}/prevclock" for i in range(n)] + ["/clk/out"]
        jit = refshim.build_jit(root, 3, True)
        rng = random.Random(width)
        rows, pokes = [], []
        for step in range(16):
            v = rng.getrandbits(width)
            pokes.append(v)
            jit.set_pin_state("/d/pin", v)
            jit.step()
            rows.append([jit.get_pin_state(p) for p in probes])
        regs.append({"name": "register_shift_w%d" % width, "kind": "register_intent", "width": width,
                     "n": n, "probes": probes, "d_values": pokes, "values": rows,
                     "note": "reference JIT with ONLY the Register emitter patched to its intent; "
                             "the unpatched emitter does not compile (simulator.py:115-127)"})
    for seed in range(8):
        made, probes = circuits.random_circuit(
            D, 1000 + seed, n_nodes=16,
            primitives=("Gate", "Not", "Constant", "Clock", "Counter", "Adder", "Register", "Register"))
        jit = refshim.build_jit(made, 3, True)
        rows = []
        for _ in range(12):
            jit.step()
            rows.append([jit.get_pin_state(p) for p in probes])
        regs.append({"name": "register_fuzz_seed%d" % seed, "kind": "register_fuzz",
                     "seed": 1000 + seed, "n_nodes": 16, "probes": probes, "values": rows,
                     "note": "intent-patched Register emitter"})
    with open(os.path.join(OUT, "register_intent.json"), "w") as f:
        json.dump(regs, f, separators=(",", ":"))
    print("register_intent.json:", len(regs), "cases")


if __name__ == "__main__":
    main()
