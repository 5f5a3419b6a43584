"""Run the persistent level kernel on a few small tiles (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits, compiler
from mcircuit_b200.engine import B200Engine
T = int(sys.argv[1]); hints = int(sys.argv[2]); mode = sys.argv[3] if len(sys.argv) > 3 else "persistent"
root = circuits.random_dag(n_gates=1_000_000, depth=200, n_inputs=4096)
eng = B200Engine(root, 1, True, lanes=64 * T * 3, keep="ports", mode=mode, tile_words=T)
eng.rt.set_tuning("persist_hints", hints); eng.rt.set_tuning("persist_ctas_per_sm", 2); eng.rt.set_tuning("persist_block", 512)
eng.randomize_inputs()
eng.run(2)
torch.cuda.synchronize()
print("ok", eng.n_tiles, eng.algorithmic_bytes(1) / eng.n_tiles)
