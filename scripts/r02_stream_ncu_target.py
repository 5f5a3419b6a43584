"""One k_run_stream launch on C4 for ncu: python scripts/r02_stream_ncu_target.py LANES STEPS"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine
lanes, steps = int(sys.argv[1]), int(sys.argv[2])
eng = B200Engine(circuits.lfsr_pipeline(), 1, True, lanes=lanes, mode="serial", keep="ports", share_edge_detectors=True)
eng.randomize_inputs(["/l%d_q%d/out" % (n, b) for n in range(256) for b in (0, 5, 17)])
eng.run(steps)
torch.cuda.synchronize()
print("ok", eng.regions_skipped())
