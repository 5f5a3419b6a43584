"""GPU sweep of the persistent (cooperative, grid-barrier) level kernel on the C5 workload."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mcircuit_b200 import circuits, compiler
from mcircuit_b200.engine import B200Engine

ap = argparse.ArgumentParser()
ap.add_argument("--gates", type=int, default=1_000_000)
ap.add_argument("--lanes", type=int, default=1 << 21)
ap.add_argument("--tiles", default="128,256,512,1024")
ap.add_argument("--unrolls", default="1,2")
ap.add_argument("--hints", default="0,1")
ap.add_argument("--inflight", default="1,2,3")
ap.add_argument("--ctas", default="1")
ap.add_argument("--blocks", default="1024")
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
root = circuits.random_dag(n_gates=a.gates, depth=200, n_inputs=4096)
prog = compiler.compile_design(compiler.flatten(root), keep="ports")
ref = B200Engine(root, 1, True, lanes=a.lanes, keep="ports", mode="levels", program=prog, tile_words=2048)
ref.randomize_inputs(); ref.run(1)
want = ref.signature()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ref.run(a.steps); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({"mode": "levels", "tile_words": ref.tile_words, "ms_per_step": ms,
                  "alg_GBps": ref.algorithmic_bytes(1) / (ms * 1e-3) / 1e9}), flush=True)
del ref
torch.cuda.empty_cache()
ref = B200Engine(root, 1, True, lanes=a.lanes, keep="ports", mode="levels_grouped", program=prog, tile_words=2048)
ref.randomize_inputs(); ref.run(1); torch.cuda.synchronize()
got = ref.signature()
e0.record(); ref.run(a.steps); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({"mode": "levels_grouped", "tile_words": ref.tile_words, "ms_per_step": ms,
                  "alg_GBps": ref.algorithmic_bytes(1) / (ms * 1e-3) / 1e9,
                  "bit_exact": bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]))}), flush=True)
del ref
torch.cuda.empty_cache()
for T, G in [(int(x), int(g)) for x in a.tiles.split(",") for g in a.inflight.split(",")]:
    eng = B200Engine(root, 1, True, lanes=a.lanes, keep="ports", mode="persistent", tile_words=T, program=prog,
                     in_flight=G)
    eng.randomize_inputs()
    for U in [int(x) for x in a.unrolls.split(",")]:
        for H in [int(x) for x in a.hints.split(",")]:
            for K, B in [(int(x), int(y)) for x in a.ctas.split(",") for y in a.blocks.split(",")]:
                eng.rt.set_tuning("persist_unroll", U); eng.rt.set_tuning("persist_hints", H)
                eng.rt.set_tuning("persist_ctas_per_sm", K); eng.rt.set_tuning("persist_block", B)
                eng.run(1); torch.cuda.synchronize()
                got = eng.signature()
                ok = bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]))
                e0.record(); eng.run(a.steps); e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.steps
                gbs = eng.algorithmic_bytes(1) / (ms * 1e-3) / 1e9
                print(json.dumps({"mode": "persistent", "tile_words": T, "in_flight": G, "unroll": U, "hints": H, "ctas_per_sm": K, "block": B,
                                  "ms_per_step": ms, "alg_GBps": gbs, "frac": gbs / 6545.0, "bit_exact": ok}), flush=True)
    del eng
    torch.cuda.empty_cache()
