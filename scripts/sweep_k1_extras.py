"""K1 extras on the C5 workload: fan-in-sorted levels and programmatic dependent launch."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mcircuit_b200 import circuits, compiler
from mcircuit_b200.engine import B200Engine
lanes = 1 << 23
root = circuits.random_dag(n_gates=1_000_000, depth=200, n_inputs=4096)
design = compiler.flatten(root)
want = None
for sort in (0, 1):
    prog = compiler.compile_design(design, keep="ports", sort_level_by_fanin=bool(sort))
    for pdl in (0, 1):
        eng = B200Engine(root, 1, True, lanes=lanes, keep="ports", mode="levels", program=prog)
        eng.rt.set_tuning("level_pdl", pdl)
        eng.randomize_inputs()
        eng.run(1); torch.cuda.synchronize()
        got = eng.signature()
        if want is None:
            want = got
        ok = bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.run(1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps({"sort_level_by_fanin": sort, "pdl": pdl, "tile_words": eng.tile_words, "streams": eng.n_streams,
                          "ms_per_step": ms, "frac": eng.algorithmic_bytes(1) / (ms * 1e-3) / 1e9 / 6545.0, "bit_exact": ok}), flush=True)
        eng.rt.set_tuning("level_pdl", 0)
        del eng
        torch.cuda.empty_cache()
