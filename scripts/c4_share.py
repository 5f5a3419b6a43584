"""C4 with and without shared clock edge detectors (resident kernel), one B200."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine
root = circuits.lfsr_pipeline()
for lanes, steps in ((1 << 20, 40), (1 << 22, 20)):
    for share in (False, True):
        eng = B200Engine(root, 1, True, lanes=lanes, mode="resident", keep="ports", share_edge_detectors=share)
        eng.randomize_inputs()
        eng.run(2); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(steps); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"config": "C4", "lanes": lanes, "share_edge_detectors": share, "micro_ops": eng.program.n_gates,
                          "ms_per_step": ms, "steps_per_s": 1e3 / ms, "gate_evals_per_s": eng.program.n_gates * lanes / (ms * 1e-3),
                          "alg_GBps": eng.algorithmic_bytes(1) / (ms * 1e-3) / 1e9}), flush=True)
        del eng
        torch.cuda.empty_cache()
