"""GPU sweep of the resident kernel (word bytes x batch size) on the C4 workload."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits, compiler
from mcircuit_b200.engine import B200Engine
root = circuits.lfsr_pipeline()
prog = compiler.compile_circuit(root, keep="ports")
for lanes, steps in ((1 << 20, 40), (1 << 16, 100), (1 << 22, 20)):
    for wb in (8,):
        for batch in (4, 8):
            for pf in (0, 1):
                eng = B200Engine(root, 1, True, lanes=lanes, mode="resident", keep="ports", program=prog,
                                 resident_word_bytes=wb, resident_batch=batch)
                eng.randomize_inputs()
                eng.rt.set_tuning('resident_prefetch', pf)
                eng.run(2); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.run(steps); e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                print(json.dumps({"lanes": lanes, "word_bytes": wb, "batch": batch, "prefetch": pf, "batches": eng.resident_batches,
                                  "ms_per_step": ms, "gate_evals_per_s": prog.n_gates * lanes / (ms * 1e-3),
                                  "alg_GBps": eng.algorithmic_bytes(1) / (ms * 1e-3) / 1e9}), flush=True)
                del eng
                torch.cuda.empty_cache()
