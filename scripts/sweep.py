"""GPU tuning sweep for the level kernel on the C5 workload (not part of the product)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mcircuit_b200 import circuits, compiler
from mcircuit_b200.engine import B200Engine

ap = argparse.ArgumentParser()
ap.add_argument("--gates", type=int, default=1_000_000)
ap.add_argument("--lanes", type=int, default=1 << 23)
ap.add_argument("--tiles", default="512,1024,2048,4096,8192")
ap.add_argument("--unrolls", default="4")
ap.add_argument("--caches", default="1")
ap.add_argument("--blocks", default="256")
ap.add_argument("--ctas", default="0")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--streams", default="1")
ap.add_argument("--hints", default="1")
a = ap.parse_args()
root = circuits.random_dag(n_gates=a.gates, depth=200, n_inputs=4096)
prog = compiler.compile_design(compiler.flatten(root), keep="ports")
res = []
for T, NS, HN in [(int(x), int(y), int(z)) for x in a.tiles.split(",") for y in a.streams.split(",") for z in a.hints.split(",")]:
    eng = B200Engine(root, 1, True, lanes=a.lanes, keep="ports", mode="levels", tile_words=T, program=prog, streams=NS,
                     level_hints=bool(HN))
    eng.randomize_inputs()
    for U in [int(x) for x in a.unrolls.split(",")]:
        for C in [int(x) for x in a.caches.split(",")]:
            for B in [int(x) for x in a.blocks.split(",")]:
                for K in [int(x) for x in a.ctas.split(",")]:
                    eng.rt.set_tuning("level_unroll", U); eng.rt.set_tuning("level_cache", C)
                    eng.rt.set_tuning("level_block", B); eng.rt.set_tuning("level_ctas_per_sm", K)
                    eng.run(1); torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(a.steps):
                        eng.run(1)          # one pass over ALL tiles per step, like bench.py
                    e1.record(); torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / a.steps
                    gbs = eng.algorithmic_bytes(1) / (ms * 1e-3) / 1e9
                    r = {"tile_words": T, "streams": NS, "hints": HN, "unroll": U, "cache": C, "block": B, "ctas_per_sm": K, "ms_per_step": ms,
                         "alg_GBps": gbs, "frac": gbs / 6545.0, "gate_evals_per_s": prog.n_gates * a.lanes / (ms * 1e-3)}
                    res.append(r); print(json.dumps(r), flush=True)
    del eng
    torch.cuda.empty_cache()
