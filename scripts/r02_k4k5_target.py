"""K4 / K5 kernels for ncu: stimulus, signature, toggle counts, pack / unpack, capture on a 20k-gate DAG at 2^22 lanes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine
root = circuits.random_dag(n_gates=20_000, depth=40, n_inputs=256, seed=77)
eng = B200Engine(root, 1, True, lanes=1 << 22, keep="ports", mode="levels")
eng.randomize_inputs()                                  # k_stimulus
eng.enable_toggle_counts(eng.output_pins()[:64])
eng.run(2)                                              # k_toggle between steps
pc, cs = eng.signature()                                # k_signature
vals = np.arange(eng.lanes, dtype=np.uint64) & np.uint64(1)
eng.set_pin_state("/i0/pin", vals)                      # k_pack
back = eng.get_pin_state("/o0/pin", lanes=True)         # k_unpack
eng.capture(eng.output_pins()[:8], 2)                   # k_capture
torch.cuda.synchronize()
print("ok", int(pc.sum()))
