"""Reference arm worker: build ONE circuit with the reference's own descriptors, compile it with the
reference's own `JIT` (core/simulator.py:174-297, loaded unmodified from baseline/_ref through the
llvmlite compatibility shim oracle/refshim.py), then time `burst()` on one core.

  python baseline/jit_worker.py <case> [--seconds S] [--wait]

Prints one JSON line when the build is done ({"built": ...}); with --wait it then blocks until a
line arrives on stdin (so the parent can keep the timing away from its own CPU-heavy phases), times
`burst()` for >= S seconds and prints the result line.  `burst()` runs burst_size + 1 steps
(simulator.py:233-242), and that is what is counted.  Single-threaded, like the reference."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MCIRCUIT_REFERENCE", os.path.join(ROOT, "baseline", "_ref"))


def make_case(name, M):
    """-> (root, gates, lanes, burst_size, description).  `gates` counts 1-bit gate evaluations per
    step per lane the way the GPU arm does (a w-bit primitive = its lowered 1-bit micro-ops)."""
    from mcircuit_b200 import circuits
    if name.startswith("c5_"):                    # c5_<gates>_<depth>_w<width>
        _, g, d, w = name.split("_")
        g, d, w = int(g), int(d), int(w[1:])
        root = circuits.random_dag(M, n_gates=g, depth=d, n_inputs=max(64, min(4096, g // 16)), width=w)
        return root, g, w, 2000, "C5 generator, %d gates, depth %d, width=%d (%d lanes per gate op)" % (g, d, w, w)
    if name == "c1":
        return circuits.counter_circuit(M), 4, 1, 501, "examples/counter.py circuit (Clock + Counter(16)), burst_size 501"
    if name.startswith("c2_w"):
        w = int(name[4:])
        return circuits.ripple_adder(M, 32, w), 160, w, 20000, "32-bit ripple adder from 5-gate full adders, width=%d" % w
    raise SystemExit("unknown case %s" % name)


def main():
    name = sys.argv[1]
    seconds = float(sys.argv[sys.argv.index("--seconds") + 1]) if "--seconds" in sys.argv else 2.0
    from oracle import refshim
    if not refshim.available():
        print(json.dumps({"case": name, "unavailable": "no reference at %s (or llvmlite / networkx missing)"
                          % refshim.REFERENCE_ROOT}), flush=True)
        return
    desc, sim = refshim.load()
    root, gates, lanes, burst_size, what = make_case(name, desc)
    t = time.perf_counter()
    with refshim.scratch_cwd():
        jit = sim.JIT(root, burst_size, True)
    build_s = time.perf_counter() - t
    print(json.dumps({"case": name, "built": True, "build_s": build_s}), flush=True)
    if "--wait" in sys.argv:
        sys.stdin.readline()
    jit.burst()                                   # warm-up
    calls, t0 = 0, time.perf_counter()
    while True:
        jit.burst()
        calls += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    steps = calls * (burst_size + 1)
    print(json.dumps({"case": name, "what": what, "gates": gates, "width": lanes, "build_s": round(build_s, 2),
                      "burst_calls": calls, "steps": steps, "seconds": dt, "us_per_step": 1e6 * dt / steps,
                      "value": gates * lanes * steps / dt, "unit": "gate-evals/s", "cores": 1,
                      "kind": "reference_jit"}), flush=True)


if __name__ == "__main__":
    main()
