"""C4 on the serial kernel: prefetch level (L1 / L2) x distance x record-line prefetch."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine
root = circuits.lfsr_pipeline()
lanes = 1 << 20
for pf in (0, 16, 32, 64, 128, 256):
    eng = B200Engine(root, 1, True, lanes=lanes, mode="serial", keep="ports", share_edge_detectors=True,
                     serial_prefetch=pf, serial_word_bytes=8)
    eng.randomize_inputs()
    for l1 in (1, 0):
        for ra in (0, 2, 4):
            if pf == 0 and l1 == 0:
                continue
            eng.rt.set_tuning("serial_pf_l1", l1)
            eng.rt.set_tuning("serial_rec_ahead", ra)
            eng.run(2); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.run(40); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 40
            print(json.dumps({"config": "C4", "lanes": lanes, "prefetch_distance": pf, "pf_l1": l1, "rec_ahead": ra,
                              "ms_per_step_avg": ms, "ms_per_rising_step_est": 2 * ms}), flush=True)
    del eng
    torch.cuda.empty_cache()
