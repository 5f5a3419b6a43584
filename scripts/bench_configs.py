"""Throughput of the other BASELINE.json configs (C1-C4) on one B200, each next to the oracle
port (oracle/mcref.c, all host threads) on a bounded sample.  Prints one JSON line per case."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine
from oracle import cref

def gpu_time(eng, steps, reps=3):
    eng.run(steps); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(steps); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best

def cpu_rate(root, lanes, steps):
    sim = cref.CRef(root, 1, True, lanes=lanes)
    sim.run(1)
    t = time.perf_counter(); sim.run(steps); dt = time.perf_counter() - t
    return dt, sim.threads

def case(name, root, lanes, steps, modes, cpu_lanes, cpu_steps, **kw):
    for mode in modes:
        eng = B200Engine(root, 1, True, lanes=lanes, mode=mode, **kw)
        eng.randomize_inputs()
        dt = gpu_time(eng, steps)
        g = eng.program.n_gates
        out = {"config": name, "mode": mode, "lanes": lanes, "steps": steps, "gates": g, "levels": eng.program.n_levels,
               "slots": eng.program.n_slots, "tiles": eng.n_tiles, "seconds": dt,
               "gate_evals_per_s": g * lanes * steps / dt, "steps_per_s": steps / dt,
               "alg_GBps": eng.algorithmic_bytes(steps) / dt / 1e9}
        print(json.dumps(out), flush=True)
        del eng
        torch.cuda.empty_cache()
    dt, th = cpu_rate(root, cpu_lanes, cpu_steps)
    print(json.dumps({"config": name, "mode": "cpu_port", "lanes": cpu_lanes, "steps": cpu_steps, "threads": th,
                      "seconds": dt, "gate_evals_per_s": g * cpu_lanes * cpu_steps / dt}), flush=True)

which = sys.argv[1:] or ["c1", "c2", "c3", "c4"]
if "c1" in which:
    case("C1 counter, single instance (64 lanes), 2,000 steps = 1,000 cycles", circuits.counter_circuit(), 64, 2000,
         ["resident"], 64, 2000)
    case("C1 counter, 2^22 instances, 2,000 steps", circuits.counter_circuit(), 1 << 22, 2000, ["resident"], 1 << 16, 2000)
if "c2" in which:
    case("C2 32-bit ripple adder, 2^20 vectors, 1 step", circuits.ripple_adder(nbits=32), 1 << 20, 1,
         ["resident", "levels"], 1 << 20, 1)
    case("C2 32-bit ripple adder, 2^24 vectors, 1 step", circuits.ripple_adder(nbits=32), 1 << 24, 1,
         ["resident", "levels"], 1 << 20, 1, keep="ports")
if "c3" in which:
    case("C3 64x64 array multiplier, 2^24 vectors, 1 step", circuits.array_multiplier(n=64), 1 << 24, 1,
         ["levels", "resident"], 1 << 18, 1, keep="ports")
if "c4" in which:
    case("C4 256 LFSR + 1k pipeline, 2^20 lanes, 200 steps (of 2e5)", circuits.lfsr_pipeline(), 1 << 20, 200,
         ["resident"], 1 << 16, 20, keep="ports")
    case("C4 256 LFSR + 1k pipeline, 2^16 lanes, 2000 steps (of 2e5)", circuits.lfsr_pipeline(), 1 << 16, 2000,
         ["resident"], 1 << 16, 20, keep="ports")
