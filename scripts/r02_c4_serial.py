"""C4 (256 LFSRs + 1,000-register pipeline) on the serial kernel: word bytes x sharing x guard x prefetch."""
import json, os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine
root = circuits.lfsr_pipeline()
quick = "--quick" in sys.argv
cases = []
for lanes in ((1 << 20,) if quick else (1 << 20, 1 << 16, 1 << 12, 1 << 22)):
    for share, guard in ((True, True), (True, False), (False, False)):
        for wb in (8, 4, 2):
            for pf in ((64,) if (quick or lanes != 1 << 20) else (64, 0, 128)):
                cases.append((lanes, share, guard, wb, pf))
for lanes, share, guard, wb, pf in cases:
    eng = B200Engine(root, 1, True, lanes=lanes, mode="serial", keep="ports", share_edge_detectors=share,
                     serial_guard=guard, serial_prefetch=pf, serial_word_bytes=wb)
    eng.randomize_inputs()
    eng.run(2); torch.cuda.synchronize()
    steps = 40 if lanes >= 1 << 20 else 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": "C4", "kernel": "k_run_serial", "lanes": lanes, "share_edge_detectors": share, "guard": guard,
                      "word_bytes": wb, "prefetch_distance": pf, "micro_ops": eng.program.n_gates,
                      "records": len(eng.program.records), "ms_per_step": ms, "steps_per_s": 1e3 / ms,
                      "gate_evals_per_s": eng.program.n_gates * lanes / (ms * 1e-3),
                      "alg_GBps": eng.algorithmic_bytes(1) / (ms * 1e-3) / 1e9, "regions_skipped_warp0": eng.regions_skipped(),
                      "info": eng.program.serial_info}), flush=True)
    del eng
    torch.cuda.empty_cache()
