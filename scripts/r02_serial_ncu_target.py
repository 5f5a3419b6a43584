"""One serial launch of C4 for ncu: python scripts/r02_serial_ncu_target.py [lanes] [steps] [--noguard]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine
args = [a for a in sys.argv[1:] if not a.startswith("--")]
lanes = int(args[0]) if args else 1 << 20
steps = int(args[1]) if len(args) > 1 else 4
eng = B200Engine(circuits.lfsr_pipeline(), 1, True, lanes=lanes, mode="serial", keep="ports", share_edge_detectors=True,
                 serial_guard="--noguard" not in sys.argv, serial_word_bytes=8)
eng.randomize_inputs()
eng.run(steps); torch.cuda.synchronize()
print("ok", eng.program.serial_info)
