"""C4 (256 LFSRs + 1,000-register pipeline) on the streamed serial kernel (K3 v3): lanes x word bytes x ring groups."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine
root = circuits.lfsr_pipeline()
quick = "--quick" in sys.argv
cases = []
for lanes in ((1 << 20,) if quick else (1 << 20, 1 << 16, 1 << 12, 1 << 22)):
    for share in (True, False):
        for wb in ((8,) if lanes >= 1 << 20 else (8, 1)):
            for groups in ((0, 4, 8) if (lanes == 1 << 20 and share) else (0,)):
                cases.append((lanes, share, wb, groups))
for lanes, share, wb, groups in cases:
    eng = B200Engine(root, 1, True, lanes=lanes, mode="serial", keep="ports", share_edge_detectors=share, serial_word_bytes=wb)
    eng.rt.set_tuning("stream_groups", groups)
    eng.randomize_inputs(["/l%d_q%d/out" % (n, b) for n in range(256) for b in (0, 5, 17)])
    eng.run(2); torch.cuda.synchronize()
    steps = 100 if lanes >= 1 << 20 else 1000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": "C4", "kernel": "k_run_stream", "lanes": lanes, "share_edge_detectors": share,
                      "word_bytes": wb, "stream_groups": groups, "micro_ops": eng.program.n_gates,
                      "ms_per_step": ms, "steps_per_s": 1e3 / ms,
                      "gate_evals_per_s": eng.program.n_gates * lanes / (ms * 1e-3),
                      "alg_GBps": eng.algorithmic_bytes(1) / (ms * 1e-3) / 1e9, "regions_skipped_cta0": eng.regions_skipped(),
                      "info": eng.program.serial_info}), flush=True)
    eng.rt.set_tuning("stream_groups", 0)
    del eng
    torch.cuda.empty_cache()
