"""CPU-only fuzz campaigns (not collected by pytest): random circuits through the compilers and the numpy test double
against the oracle, with every option, execution mode, producer model and launch split drawn at random.

  python scripts/fuzz_cpu.py stream 1 400      # stream programs (K3): any circuit, any optimisation subset
  python scripts/fuzz_cpu.py auto 11 300       # mode='auto' / explicit modes, prevclock pokes (un-sharing), cross-mode checkpoints
  python scripts/fuzz_cpu.py chains 21 250     # register rows + pipelines (C4's shape): regions, patterns, stash, by-record stages

Round 2: 6 x 400 + 4 x 300 + 6 x 250 cases, 0 failures."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import mcircuit_b200.core.descriptors as MD  # noqa: E402
from _cpu_runtime import CpuRuntime  # noqa: E402
from mcircuit_b200 import circuits, compiler  # noqa: E402
from mcircuit_b200.engine import B200Engine  # noqa: E402
from oracle import cref, restate  # noqa: E402

ALL = ("Gate", "Not", "Constant", "Clock", "Counter", "Adder", "Register", "Register", "Counter")


def stream_campaign(seed, count):
    rnd = random.Random(seed)
    n_fail = 0
    for it in range(count):
        seed = rnd.randrange(10**9)
        kw = dict(seed=seed, n_nodes=rnd.randint(2, 60), max_width=rnd.choice([1, 2, 5, 8, 33, 64]), feedback=rnd.random() < 0.7,
                  hierarchy=rnd.random() < 0.5, primitives=ALL)
        opts = dict(patterns=rnd.random() < 0.8, stash=rnd.random() < 0.8, ring=rnd.random() < 0.8)
        share, guard, prefix, map_pins = rnd.random() < 0.6, rnd.random() < 0.8, rnd.random() < 0.3, rnd.random() < 0.7
        keep = rnd.choice(["all", "ports"]); model = rnd.choice(["early", "late"]); cta = rnd.choice([1, 2, 128])
        chunks = [rnd.randint(1, 50) for _ in range(rnd.randint(1, 4))]
        try:
            root, probes = circuits.random_circuit(MD, **kw)
            lanes = 64 * 3
            rt = CpuRuntime(); rt.serial_fetch_model, rt.serial_cta_cols = model, cta
            ora = restate.RestatedJIT(root, 2, map_pins, lanes=lanes)
            eng = B200Engine(root, 2, map_pins, lanes=lanes, runtime=rt, keep=keep, mode="serial", share_edge_detectors=share,
                             serial_guard=guard, adder_prefix=prefix, serial_options=opts)
            g = np.random.default_rng(seed)
            d = eng.program.design
            for p in ("/in0/pin", "/in1/pin"):
                w = d.pin_width[d.find_pin(p)]
                v = g.integers(0, (1 << w) - 1, size=lanes, dtype=np.uint64, endpoint=True)
                ora.set_pin_state(p, v); eng.set_pin_state(p, v)
            for n in chunks:
                eng.run(n)
                for _ in range(n): ora.step()
                for p in probes:
                    try: got = eng.get_pin_state(p, lanes=True)
                    except compiler.PinNotObservable: continue
                    assert np.array_equal(got, ora.get_pin_state(p, lanes=True)), (n, p)
        except Exception as exc:
            n_fail += 1
            print("FAIL", kw, opts, share, guard, prefix, map_pins, keep, model, cta, chunks, repr(exc)[:200], flush=True)
    print("done:", count, "cases,", n_fail, "failures", flush=True)
    return n_fail


def auto_campaign(seed, count):
    rnd = random.Random(seed)
    n_fail = 0
    for it in range(count):
        seed = rnd.randrange(10**9)
        kw = dict(seed=seed, n_nodes=rnd.randint(2, 50), max_width=rnd.choice([1, 2, 5, 8, 33, 64]), feedback=rnd.random() < 0.7,
                  hierarchy=rnd.random() < 0.5, primitives=ALL)
        mode = rnd.choice(["auto", "auto", "levels", "resident", "serial"])
        keep = rnd.choice(["auto", "all", "ports"]); prefix = rnd.random() < 0.4; map_pins = rnd.random() < 0.7
        try:
            root, probes = circuits.random_circuit(MD, **kw)
            lanes = 64 * rnd.choice([1, 3])
            ora = restate.RestatedJIT(root, 2, map_pins, lanes=lanes)
            eng = B200Engine(root, 2, map_pins, lanes=lanes, runtime=CpuRuntime(), keep=keep, mode=mode, adder_prefix=prefix)
            g = np.random.default_rng(seed)
            d = eng.program.design
            pins = [p[0] for p in d.iter_pins()]
            prev = [p for p in pins if p.endswith("/prevclock")]
            for p in ("/in0/pin", "/in1/pin"):
                w = d.pin_width[d.find_pin(p)]
                v = g.integers(0, (1 << w) - 1, size=lanes, dtype=np.uint64, endpoint=True)
                ora.set_pin_state(p, v); eng.set_pin_state(p, v)
            snap = None
            for phase in range(4):
                n = rnd.randint(1, 12)
                eng.run(n)
                for _ in range(n): ora.step()
                if phase == 1 and prev and rnd.random() < 0.5:            # poke one member's prevclock (un-shares if shared)
                    p = rnd.choice(prev); v = g.integers(0, 2, size=lanes, dtype=np.uint64)
                    ora.set_pin_state(p, v); eng.set_pin_state(p, v)
                if phase == 2 and rnd.random() < 0.5:                      # checkpoint into a fresh engine of another mode
                    other = B200Engine(root, 2, map_pins, lanes=lanes, runtime=CpuRuntime(), keep=keep,
                                       mode=rnd.choice(["levels", "resident", "serial"]), adder_prefix=not prefix,
                                       share_edge_detectors=rnd.random() < 0.5)
                    other.load_state_dict(eng.state_dict()); eng = other
                for p in probes:
                    try: got = eng.get_pin_state(p, lanes=True)
                    except compiler.PinNotObservable: continue
                    assert np.array_equal(got, ora.get_pin_state(p, lanes=True)), (phase, p)
        except Exception as exc:
            n_fail += 1
            print("FAIL", kw, mode, keep, prefix, map_pins, repr(exc)[:300], flush=True)
    print("done:", count, "cases,", n_fail, "failures", flush=True)
    return n_fail


def register_chain_campaign(seed, count):
    rnd = random.Random(seed)
    n_fail = 0
    for it in range(count):
        bits = rnd.choice([3, 5, 8, 16, 31, 32, 40])
        n_l = rnd.randint(1, 12); stages = rnd.choice([0, 1, 7, 33, 100, 150])
        taps = tuple(sorted({rnd.randrange(bits) for _ in range(rnd.randint(1, 4))}, reverse=True))
        couple = rnd.random() < 0.8
        opts = dict(patterns=rnd.random() < 0.85, stash=rnd.random() < 0.85, ring=rnd.random() < 0.85)
        share, guard = rnd.random() < 0.8, rnd.random() < 0.85
        keep = rnd.choice(["all", "ports"]); model = rnd.choice(["early", "late"]); cta = rnd.choice([1, 2, 128])
        chunks = [rnd.randint(1, 40) for _ in range(rnd.randint(1, 4))]
        desc = (n_l, bits, stages, taps, couple, opts, share, guard, keep, model, cta, chunks)
        try:
            root = circuits.lfsr_pipeline(MD, n_l, bits, stages, taps=taps, couple=couple)
            lanes = 64 * 3
            rt = CpuRuntime(); rt.serial_fetch_model, rt.serial_cta_cols = model, cta
            ora = cref.CRef(root, 1, True, lanes=lanes)
            eng = B200Engine(root, 1, True, lanes=lanes, runtime=rt, keep=keep, mode="serial", share_edge_detectors=share,
                             serial_guard=guard, serial_options=opts)
            g = np.random.default_rng(it)
            for n in range(n_l):
                for b in range(bits):
                    if g.random() < 0.5:
                        v = g.integers(0, 2, size=lanes, dtype=np.uint64)
                        eng.set_pin_state(f"/l{n}_q{b}/out", v); ora.set_pin_state(f"/l{n}_q{b}/out", v)
            probes = eng.output_pins() + ([f"/l0_q{bits-1}/out", "/l0_q0/prevclock"] if keep == "all" else [])
            for n in chunks:
                eng.run(n); ora.run(n)
                for p in probes:
                    assert np.array_equal(eng.get_pin_planes(p), ora.planes(p)), (n, p)
        except Exception as exc:
            n_fail += 1
            print("FAIL", desc, repr(exc)[:300], flush=True)
    print("done:", count, "cases,", n_fail, "failures", flush=True)
    return n_fail


if __name__ == "__main__":
    which, seed, count = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    fn = {"stream": stream_campaign, "auto": auto_campaign, "chains": register_chain_campaign}[which]
    sys.exit(1 if fn(seed, count) else 0)
