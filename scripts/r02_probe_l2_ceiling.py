"""What does the level kernel K1 reach when every fan-in row is an L2 hit?  Same kernel, same
launch path (cached step graphs, PDL, 2 streams) as the C5 bench, but in1 is drawn from the last
`window` levels only, so the live set of a 512-word tile (window+1 levels x 5,000 rows x 4 KB)
stays inside the 126 MB L2.  The result is the L2-side ceiling of K1; C5's in1 rows are random
over ALL earlier levels and mostly miss."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine

lanes = 1 << 22
for window, tile in ((1, 512), (1, 256), (2, 512), (1, 1024), (None, 512)):
    root = circuits.random_dag(n_gates=200_000, depth=40, n_inputs=4096, in1_window=window)
    eng = B200Engine(root, 1, True, lanes=lanes, mode="levels", keep="ports", tile_words=tile)
    eng.randomize_inputs()
    eng.run(1); eng.run(1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run(5); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"probe": "k1_l2_ceiling", "in1_window_levels": window, "tile_words": tile, "lanes": lanes,
                      "scratch_slots": eng.program.n_scratch, "persist_slots": eng.program.n_persist,
                      "live_MB_per_tile": (eng.program.n_scratch) * tile * 8 / 1e6,
                      "ms_per_step": ms, "alg_GBps": eng.algorithmic_bytes(1) / (ms * 1e-3) / 1e9}), flush=True)
    del eng
    torch.cuda.empty_cache()
