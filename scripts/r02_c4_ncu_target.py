"""One resident launch of C4 (2^20 lanes, 4 steps) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine
share = "--share" in sys.argv
lanes = 1 << 20
eng = B200Engine(circuits.lfsr_pipeline(), 1, True, lanes=lanes, mode="resident", keep="ports", share_edge_detectors=share)
eng.randomize_inputs()
eng.run(4); torch.cuda.synchronize()
print("ok", eng.program.n_gates, eng.program.n_levels)
