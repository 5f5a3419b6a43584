#!/usr/bin/env python
"""bench.py - gate-evals/sec of the levelized evaluation hot path (BASELINE.json metric).

Workload (N GPUs, weak scaling): BASELINE.json configs[4] - synthetic random levelized DAG,
1M two-input gates, depth 200, 4096 primary inputs - on 2^23 stimulus lanes PER GPU (2^26
lanes at 8 GPUs, the named configuration), netlist replicated, lanes sharded, one
NCCL all-reduce (sum) of the per-GPU output signatures per step.  One "step" = one
evaluation pass of the whole netlist over every lane of the rank's shard + its signature.

  python bench.py [--gpus N --steps K --warmup W]            our engine (CUDA, sm_100a)
  python bench.py --impl reference [...]                      the reference algorithm on the host
                                                              cores (oracle port, all threads),
                                                              bounded sample of the same workload

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gate_evals_per_sec"
UNIT = "gate-evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gates", type=int, default=1_000_000)
    ap.add_argument("--depth", type=int, default=200)
    ap.add_argument("--inputs", type=int, default=4096)
    ap.add_argument("--lanes-per-gpu", type=int, default=1 << 23)
    ap.add_argument("--tile-words", type=int, default=0, help="0 = engine default")
    ap.add_argument("--mode", default="levels", choices=["levels", "resident", "auto"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU baseline duration")
    ap.add_argument("--streams", type=int, default=0, help="compute streams that independent tiles alternate over (0 = engine default)")
    ap.add_argument("--tune", default="", help="comma list key=value for mcb_set_tuning")
    ap.add_argument("--verify", action="store_true", help="check sampled columns against the oracle")
    return ap.parse_args()


def workload_name(a):
    return "c5_random_dag_%dk_gates_depth%d_%d_lanes_per_gpu" % (a.gates // 1000, a.depth, a.lanes_per_gpu)


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except Exception:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port (plain C restatement of the reference's
# sequential in-place evaluation) on the host cores, bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_baseline(a, root, seconds, rounds=1, mem_bytes=6 << 30):
    """Time oracle/mcref.c (the reference's sequential in-place pass, many lanes wide) on all
    host threads.  The oracle keeps one row per NET for every lane word (the reference's model:
    one global per net, nothing recycled), so the sample's lanes are bounded by memory and the
    sample's duration is reached by repeating full passes over it."""
    import numpy as np
    from oracle import cref, restate
    threads = os.cpu_count() or 1
    quantum = 8 * threads                                 # one 8-column block per thread
    n_rows = a.gates + a.inputs
    words = max(quantum, min(1 << 14, mem_bytes // (n_rows * 8)) // quantum * quantum)
    sim = cref.CRef(root, 1, True, lanes=64 * words, threads=threads)
    sim.set_rows(sim._base("/i0/pin"),
                 restate.stimulus_word(np.arange(a.inputs)[:, None], np.arange(sim.words)[None, :]))
    sim.run(1)                                            # warm-up (first touch of the net file)
    t = time.perf_counter()
    sim.run(1)                                            # calibration
    dt = time.perf_counter() - t
    passes = int(max(1, min(200, round(seconds / max(dt, 1e-6)))))
    times = []
    for _ in range(rounds):
        t = time.perf_counter()
        sim.run(passes)
        times.append(time.perf_counter() - t)
    best = min(times)
    value = a.gates * 64 * words * passes / best
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "full %d-gate netlist, %d lanes (%d of the %d lane words of one GPU shard), %d "
                      "passes, oracle/mcref.c on %d threads, %.1f s" % (
                          a.gates, 64 * words, words, a.lanes_per_gpu // 64, passes, threads, best),
            "seconds": best, "times": times, "sim": sim, "passes": passes}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mcircuit_b200 import circuits
    root = circuits.random_dag(n_gates=a.gates, depth=a.depth, n_inputs=a.inputs)
    per_step = max(1.0, min(15.0, 100.0 / max(1, a.steps + a.warmup)))
    cb = cpu_baseline(a, root, per_step, rounds=1)
    sim, passes = cb.pop("sim"), cb.pop("passes")
    times = []
    for i in range(a.warmup + a.steps):                   # one step = `passes` passes over the sample
        t = time.perf_counter()
        sim.run(passes)
        dt = time.perf_counter() - t
        if i >= a.warmup:
            times.append(dt)
    total = sum(times)
    lanes = sim.lanes
    value = a.gates * lanes * passes * len(times) / total
    cb["value"] = value
    cb.pop("times", None)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "gates": a.gates, "depth": a.depth, "inputs": a.inputs,
                   "sample_lanes": lanes, "passes_per_step": passes, "host_threads": cb["cores"],
                   "note": "reference algorithm (sequential in-place emit-order pass, "
                           "core/simulator.py:195-223) as the oracle's plain-C port; the reference's "
                           "own LLVM JIT cannot be built at 1M gates (SURVEY Q13) and is Python, so it "
                           "does not travel to this box"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """Pin this rank's threads (and therefore its pinned host buffers, first touch) to the
    NUMA node its GPU hangs off, so host<->device copies do not cross the socket link."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from mcircuit_b200 import circuits
    from mcircuit_b200.engine import B200Engine

    t_build = time.time()
    root = circuits.random_dag(n_gates=a.gates, depth=a.depth, n_inputs=a.inputs)
    words = a.lanes_per_gpu // 64
    eng = B200Engine(root, 1, True, lanes=a.lanes_per_gpu, device=local, keep="ports", mode=a.mode,
                     tile_words=(a.tile_words or None), lane_word_offset=rank * words, streams=(a.streams or None))
    t_build = time.time() - t_build
    for kv in filter(None, a.tune.split(",")):
        k, v = kv.split("=")
        eng.rt.set_tuning(k, int(v))
    prog = eng.program
    eng.randomize_inputs()                       # stimulus generated on the device (K5)
    sig = eng.signature_plan()
    in_plan = eng.io_plan(eng.input_pins())
    out_plan = eng.io_plan(eng.output_pins())

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step(ev=None):
        if ev is not None:
            ev[0].record()
        eng.run(1)
        if ev is not None:
            ev[1].record()
        buf = eng.signature_run(sig)
        if world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)     # int64 wrap-around == uint64 sums
        return buf

    for _ in range(a.warmup):
        one_step()
    sampler = ClockSampler(local) if rank == 0 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    eng.rt.launch_count(reset=True)
    t0 = time.time()
    e0.record()
    for i in range(a.steps):
        buf = one_step(evs[i])
    e1.record()
    sync_all()
    t1 = time.time()
    launches = eng.rt.launch_count()
    ms = e0.elapsed_time(e1)
    eval_ms = [x.elapsed_time(y) for x, y in evs]
    if world > 1:
        t = torch.tensor([ms, sum(eval_ms)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, eval_total = float(t[0]), float(t[1])
    else:
        eval_total = sum(eval_ms)
    clocks = sampler.window(t0, t1) if sampler else None
    signature_host = eng.rt.to_host_u64(buf)

    total_lanes = a.lanes_per_gpu * world
    value = prog.n_gates * total_lanes * a.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (k_eval_level), measured live ------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    level_launches = prog.n_levels * eng.n_tiles            # per step per GPU
    alg_bytes_step = eng.algorithmic_bytes(1)               # this GPU, one step
    avg_launch_s = (eval_total * 1e-3) / (a.steps * level_launches)
    achieved = (alg_bytes_step / level_launches) / avg_launch_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tp):
        try:
            traffic = json.load(open(tp)).get("k_eval_level_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_eval_level", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_step / level_launches,
                "avg_launch_us": avg_launch_s * 1e6, "launches_per_step": level_launches,
                "eval_share_of_step": eval_total / ms,
                "note": "achieved counts ALGORITHMIC bytes (24 B per gate-word, SURVEY 8d); with %d-word tiles "
                        "part of them is served by L2 (previous level's rows, recycled scratch slots), so frac "
                        "can exceed 1 while real DRAM traffic stays below the peak - see traffic and "
                        "profiles/r01_ncu_eval_level.md" % eng.tile_words}

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region ----------------------
    e2e = None
    if not a.no_e2e:
        n_in, n_out = in_plan["n"], out_plan["n"]
        h_in = torch.empty((n_in, words), dtype=torch.int64, pin_memory=True)
        h_out = torch.empty((n_out, words), dtype=torch.int64, pin_memory=True)
        h_sig = torch.empty(sig["buf"].shape, dtype=torch.int64, pin_memory=True)
        eng.download_rows(in_plan, h_in, non_blocking=False)         # the stimulus, now on the host
        torch.cuda.synchronize(dev)

        def e2e_step():
            # H2D of every stimulus plane, evaluation and D2H of every output plane, pipelined
            # tile by tile on three streams; then the signature (+ all-reduce) and its D2H
            eng.run_host(in_plan, h_in, out_plan, h_out, steps=1)
            b = eng.signature_run(sig)
            if world > 1:
                dist.all_reduce(b, op=dist.ReduceOp.SUM)
            h_sig.copy_(b, non_blocking=True)
        for _ in range(max(1, min(a.warmup, 2))):
            e2e_step()
        sync_all()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(a.steps):
            e2e_step()
        f1.record()
        sync_all()
        ems = f0.elapsed_time(f1)
        if world > 1:
            t = torch.tensor([ems], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t[0])
        e2e = {"value": prog.n_gates * total_lanes * a.steps / (ems * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h_in.numel() * 8), "d2h_bytes_per_step": int((h_out.numel() + h_sig.numel()) * 8),
               "ms_per_step": ems / a.steps,
               "api": "B200Engine.run_host(h_in, h_out) [tile-pipelined H2D / eval / D2H over the C ABI] "
                      "-> signature -> D2H (pinned host buffers)"}
        assert np.array_equal(h_sig.numpy().view(np.uint64), signature_host), "e2e signature differs from resident run"
        # the downloaded planes really are the outputs: their popcounts are the signature
        local_sig = eng.rt.to_host_u64(eng.signature_run(sig))       # this rank's shard, not reduced
        idx = list(range(0, n_out, max(1, n_out // 16)))
        for i in idx:
            pc = int(np.unpackbits(h_out[i].numpy().view(np.uint8)).sum())
            assert pc == int(local_sig[i]), "downloaded output plane %d does not match its signature" % i
        e2e["downloaded_planes_checked_against_signature"] = len(idx)

    # ---- optional parity spot-check against the oracle on sampled columns -----------------------------
    verify = None
    if a.verify and rank == 0:
        from oracle import cref, restate
        cols = np.array([0, 1, 7, words // 2, words - 1], dtype=np.int64)
        sim = cref.CRef(root, 1, True, lanes=64 * 8, threads=1)
        stim = restate.stimulus_word(np.arange(a.inputs)[:, None], (cols + rank * words)[None, :])
        full = np.zeros((a.inputs, 8), dtype=np.uint64)
        full[:, :len(cols)] = stim
        sim.set_rows(sim._base("/i0/pin"), full)
        sim.run(1)
        want = sim.rows(sim._base("/o0/pin"), 1)
        ok = True
        for k, p in enumerate(eng.output_pins()):
            got = eng.get_pin_planes(p)[0][cols]
            if not np.array_equal(got, sim.planes(p)[0][:len(cols)]):
                ok = False
                break
        verify = {"columns": cols.tolist(), "outputs_checked": len(eng.output_pins()), "bit_exact": ok}

    cb = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cb = cpu_baseline(a, root, a.cpu_seconds)
        for k in ("sim", "times", "passes"):
            cb.pop(k, None)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(a), "gates": prog.n_gates, "depth": a.depth,
                       "inputs": a.inputs, "lanes_per_gpu": a.lanes_per_gpu, "lanes_total": total_lanes,
                       "levels": prog.n_levels, "tile_words": eng.tile_words, "tiles": eng.n_tiles,
                       "persistent_slots": prog.n_persist, "scratch_slots": prog.n_scratch,
                       "mode": eng.mode, "streams": eng.n_streams, "parallelism": "lanes sharded x%d, netlist replicated" % world,
                       "l2": "inputs larger than L2: %.1f GB of state streamed per step"
                             % (alg_bytes_step / 1e9),
                       "stimulus": "splitmix64 counter-based, generated on device",
                       "compile_s": round(t_build, 1), "tuning": a.tune, "numa_node": numa},
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "signature_popcount_sum": int(signature_host[:sig["n"]].sum()),
        }
        if verify is not None:
            out["verify"] = verify
        print(json.dumps(out))
    if sampler:
        sampler.stop()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
