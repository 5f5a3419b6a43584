"""Which execution mode for which design?  ms per step of levels / resident / serial (and what mode='auto' picks) on
designs of different shapes, 2^16 and 2^20 lanes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine
designs = [
    ("gate_level_lfsr32_x64 (1.6k NAND-level gates, feedback, no primitives)", lambda: circuits.gate_level_lfsr(n_bits=32, taps=(31, 21, 1, 0), stages=64)),
    ("raw_counter 8 bit (gate-level adders + master-slave flip-flops)", lambda: circuits.raw_counter(bits=8, seed=1)),
    ("lfsr_pipeline 32 x 32 + 200 stages (Register primitives)", lambda: circuits.lfsr_pipeline(n_lfsr=32, lfsr_bits=32, stages=200)),
    ("ripple_adder 32 (combinational, 65 levels)", lambda: circuits.ripple_adder(nbits=32)),
    ("array_multiplier 16x16 (1.4k gates, 88 levels)", lambda: circuits.array_multiplier(n=16)),
    ("random_dag 20k gates depth 40", lambda: circuits.random_dag(n_gates=20_000, depth=40, n_inputs=256, seed=77)),
]
for name, make in designs:
    for lanes in (1 << 16, 1 << 20):
        row = {"design": name, "lanes": lanes}
        for mode in ("levels", "resident", "serial", "auto"):
            try:
                eng = B200Engine(make(), 1, True, lanes=lanes, mode=mode, keep="ports")
                eng.randomize_inputs()
                steps = 50
                eng.run(4); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.run(steps); e1.record(); torch.cuda.synchronize()
                row[mode] = round(e0.elapsed_time(e1) / steps, 4)
                if mode == "auto":
                    row["auto_mode"] = eng.mode
                row["micro_ops"] = eng.program.n_gates
                del eng
                torch.cuda.empty_cache()
            except Exception as exc:
                row[mode] = "error: %r" % (exc,)
        print(json.dumps(row), flush=True)
