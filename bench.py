#!/usr/bin/env python
"""bench.py - gate-evals/sec of the levelized evaluation hot path (BASELINE.json metric).

Headline workload (N GPUs, weak scaling): BASELINE.json configs[4] - synthetic random levelized
DAG, 1M two-input gates, depth 200, 4096 primary inputs - on 2^23 stimulus lanes PER GPU (2^26
lanes at 8 GPUs, the named configuration), netlist replicated, lanes sharded, one NCCL
all-reduce (sum) of the per-GPU output signatures per step.  One "step" = one evaluation pass of
the whole netlist over every lane of the rank's shard + its signature.

  python bench.py [--gpus N --steps K --warmup W]            our engine (CUDA, sm_100a)
  python bench.py --impl reference [...]                      the reference's own LLVM JIT
                                                              (baseline/_ref, 1 core) + the oracle's
                                                              C port on all host threads

Prints ONE JSON line (rank 0).  Besides the base contract it carries
  verify        default-on parity spot check of THIS run's outputs against the oracle (rc != 0 if false)
  roofline      the level kernel K1 against the measured HBM copy peak
  cpu_baseline  port = oracle/mcref.c on all host threads; jit = the reference's JIT.burst(), 1 core
  configs       BASELINE.json configs C1-C4, measured in the same process after the headline
See DESIGN.md "Measurement" for every field.
"""
import argparse
import hashlib
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
JIT_CASES_C5 = ("c5_1000_20_w1", "c5_1000_20_w64", "c5_10000_50_w1", "c5_10000_50_w64")
JIT_CASES_SMALL = ("c1", "c2_w1", "c2_w64")


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
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle spot check (it is outside the timed region)")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs C1-C4")
    ap.add_argument("--no-jit", action="store_true", help="do not time the reference JIT beside the run")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="target duration of the C-port baseline")
    ap.add_argument("--jit-seconds", type=float, default=2.0, help="burst() timing window per JIT case")
    ap.add_argument("--c4-steps", type=int, default=200_000, help="steps of config C4 in the configs array (200000 = the named 10^5 clock cycles)")
    ap.add_argument("--streams", type=int, default=0, help="compute streams that independent tiles alternate over (0 = engine default)")
    ap.add_argument("--tune", default="", help="comma list key=value for mcb_set_tuning")
    ap.add_argument("--verify", action="store_true", help=argparse.SUPPRESS)     # round-1 flag; verification is default-on now
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
# the reference's own JIT, timed beside the run (one core per case, the reference is single-threaded)
# ------------------------------------------------------------------------------------------------
class JitWorkers:
    """One `baseline/jit_worker.py` process per case.  They build right away (the 10^4-gate builds
    take over a minute of Python flattening + LLVM codegen each) and then WAIT; `release()` lets
    them time `burst()` - called when this process is not hammering the host cores itself."""

    def __init__(self, cases, seconds):
        self.procs = {}
        worker = os.path.join(ROOT, "baseline", "jit_worker.py")
        for c in cases:
            try:
                self.procs[c] = subprocess.Popen([sys.executable, worker, c, "--seconds", str(seconds), "--wait"],
                                                 stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                                 stderr=subprocess.DEVNULL, text=True)
            except Exception:
                pass

    def release(self):
        for p in self.procs.values():
            try:
                p.stdin.write("go\n")
                p.stdin.flush()
            except Exception:
                pass

    def collect(self, timeout=420.0):
        out, deadline = {}, time.time() + timeout
        for c, p in self.procs.items():
            try:
                stdout, _ = p.communicate(timeout=max(1.0, deadline - time.time()))
                lines = [json.loads(l) for l in stdout.splitlines() if l.startswith("{")]
                out[c] = lines[-1] if lines else {"case": c, "unavailable": "worker printed nothing"}
            except subprocess.TimeoutExpired:
                p.kill()
                out[c] = {"case": c, "unavailable": "JIT build did not finish within the bench's time limit"}
            except Exception as exc:
                out[c] = {"case": c, "unavailable": repr(exc)}
        return out


def jit_note():
    return ("the reference's LLVM JIT (baseline/_ref/core/simulator.py:174-297, unmodified, llvmlite shim), burst() timed on ONE "
            "core - the reference has no threads; width=1 is its ordinary single-instance mode, width=64 its best native mode "
            "(64 lanes per gate op).  A 10^6-gate JIT cannot be built (SURVEY Q13: 5x10^4 gates = 284 s of LLVM codegen, "
            "2x10^5 did not finish in 25 min), so the same generator is timed at 10^3 and 10^4 gates")


# ------------------------------------------------------------------------------------------------
# CPU port: the oracle's plain-C restatement of the reference's sequential in-place pass
# ------------------------------------------------------------------------------------------------
def make_port(a, root, mem_bytes=6 << 30):
    """oracle/mcref.c on the full netlist.  The oracle keeps one row per NET for every lane word (the
    reference's model: one global per net, nothing recycled), so the sample's lanes are bounded by
    memory and its duration is reached by repeating full passes."""
    from oracle import cref
    threads = os.cpu_count() or 1
    quantum = 8 * threads                                 # one 8-column block per thread
    n_rows = a.gates + a.inputs
    words = max(quantum, min(1 << 14, mem_bytes // (n_rows * 8)) // quantum * quantum)
    return cref.CRef(root, 1, True, lanes=64 * words, threads=threads)


def port_stimulus(sim, a, cols):
    """Row k of the inputs = stream k of the counter-based stimulus at global columns `cols`."""
    import numpy as np
    from oracle import restate
    full = np.zeros((a.inputs, sim.words), dtype=np.uint64)
    cols = np.asarray(cols, dtype=np.int64)
    full[:, :len(cols)] = restate.stimulus_word(np.arange(a.inputs)[:, None], cols[None, :])
    sim.set_rows(sim._base("/i0/pin"), full)


def time_port(sim, a, seconds):
    sim.run(1)                                            # warm-up (first touch of the net file)
    t = time.perf_counter()
    sim.run(1)                                            # calibration
    dt = time.perf_counter() - t
    passes = int(max(1, min(200, round(seconds / max(dt, 1e-6)))))
    t = time.perf_counter()
    sim.run(passes)
    best = time.perf_counter() - t
    value = a.gates * sim.lanes * passes / best
    return {"value": value, "unit": UNIT, "cores": sim.threads, "kind": "port",
            "value_per_core": value / sim.threads, "host_cpus": os.cpu_count(),
            "sample": "full %d-gate netlist, %d lanes (%d of the %d lane words of one GPU shard), %d passes, "
                      "oracle/mcref.c on %d threads, %.1f s" % (a.gates, sim.lanes, sim.words, a.lanes_per_gpu // 64,
                                                                 passes, sim.threads, best),
            "seconds": best, "same_config": False,
            "same_config_note": "same netlist, bounded lane sample (a rate metric: gate-evals/s does not depend on the lane count "
                                "once every thread has work)"}, passes


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from baseline import install_ref
    install_ref.install()
    # headline of this arm: the reference's JIT at the largest size of the C5 generator it can build
    jit = JitWorkers(JIT_CASES_C5 if not a.no_jit else (), max(1.0, min(a.jit_seconds, 60.0 / max(1, a.steps + a.warmup))))
    from mcircuit_b200 import circuits
    root = circuits.random_dag(n_gates=a.gates, depth=a.depth, n_inputs=a.inputs)
    sim = make_port(a, root)
    port_stimulus(sim, a, range(sim.words))
    per_step = max(1.0, min(10.0, 60.0 / max(1, a.steps + a.warmup)))
    port, passes = time_port(sim, a, per_step)
    times = []
    for i in range(a.warmup + a.steps):                   # one step = `passes` passes over the sample
        t = time.perf_counter()
        sim.run(passes)
        dt = time.perf_counter() - t
        if i >= a.warmup:
            times.append(dt)
    total = sum(times)
    port["value"] = a.gates * sim.lanes * passes * len(times) / total
    port["value_per_core"] = port["value"] / sim.threads
    jit.release()
    jits = jit.collect()
    best = None
    for c in ("c5_10000_50_w64", "c5_1000_20_w64"):
        if c in jits and "value" in jits[c]:
            best = jits[c]
            break
    if best is not None:
        # "reference": the reference's own implementation (its LLVM JIT from baseline/_ref), not a port
        value, kind, ms = best["value"], "reference", 1e3 * best["seconds"] / max(1, best["burst_calls"])
        sample = ("JIT.burst() of the C5 generator at %d gates, width=%d, %d steps in %.2f s on 1 core (the reference "
                  "cannot build 10^6 gates)" % (best["gates"], best["width"], best["steps"], best["seconds"]))
        cores = 1
    else:
        value, kind, ms, sample, cores = port["value"], "port", 1e3 * total / len(times), port["sample"], port["cores"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "gates": a.gates, "depth": a.depth, "inputs": a.inputs,
                   "same_config": False,
                   "note": "value = the reference's own LLVM JIT (one core) on the largest instance of the same generator it can "
                           "build; cpu_baseline.port = the reference algorithm (sequential in-place emit-order pass, "
                           "core/simulator.py:195-223) as the oracle's plain-C port on ALL host threads on the full 1M-gate "
                           "netlist - a stronger baseline than the JIT in absolute terms"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "what": "reference_jit" if kind == "reference" else "oracle/mcref.c", "sample": sample,
                         "host_cpus": os.cpu_count(), "jit": list(jits.values()), "jit_note": jit_note(), "port": port},
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


def build_info():
    """What binary is being timed: the sidecar written by mcircuit_b200/build.py says which source
    hash the shipped .so was compiled from; it must equal the hash of the tracked source."""
    from mcircuit_b200 import build as _b
    info = _b.read_buildinfo() or {}
    try:
        so_sha = hashlib.sha256(open(_b.OUT, "rb").read()).hexdigest()[:16]
    except Exception:
        so_sha = None
    return {"source_sha16": _b.source_sha(), "built_from_source_sha16": info.get("source_sha16"),
            "so_sha16": so_sha, "so_sha16_at_build": info.get("so_sha16"),
            "matches_tracked_source": info.get("source_sha16") == _b.source_sha() and so_sha == info.get("so_sha16"),
            "rebuilt_in_this_run": bool(getattr(_b, "rebuilt_in_this_process", False)),
            "nvcc": info.get("nvcc"), "flags": info.get("flags"), "built_on": info.get("when")}


def verify_c5(a, eng, rank, words, sim=None):
    """Parity of THIS run: sampled lane columns of this rank's shard, spread over every compute
    stream's tiles (first / middle / last tile, both planes), all outputs, bit for bit against the
    oracle's C restatement fed the same counter-based stimulus."""
    import numpy as np
    from oracle import cref
    tw, nt = eng.tile_words, eng.n_tiles
    picks = [0, 1, tw - 1, tw, 2 * tw - 1 if nt > 1 else words // 3, (nt // 2) * tw + 7, (nt // 2 + 1) * tw + tw // 2,
             words - tw - 1, words - 2, words - 1]
    cols = np.array(sorted({int(c) % words for c in picks}), dtype=np.int64)
    if sim is None:
        sim = cref.CRef(eng.root, 1, True, lanes=64 * max(8, (len(cols) + 7) // 8 * 8), threads=1)
    port_stimulus(sim, a, cols + rank * words)
    sim.run(1)
    pins = eng.output_pins()
    want = sim.rows(sim._base(pins[0]), len(pins))[:, :len(cols)]
    got = eng.sample_planes(pins, cols)
    ok = bool(np.array_equal(got, want))
    return {"bit_exact": ok, "columns": cols.tolist(), "tiles_touched": sorted({int(c) // tw for c in cols}),
            "streams_touched": sorted({(int(c) // tw) % eng.n_streams for c in cols}),
            "outputs_checked": len(pins), "words_compared": int(got.size),
            "against": "oracle/mcref.c (C restatement of core/simulator.py:195-223), same splitmix64 stimulus"}, sim


def run_configs(a, dev, world, rank, jits):
    """BASELINE.json configs C1-C4 on this GPU, each checked against the oracle and timed with CUDA
    events; at N > 1 every rank runs the same per-GPU shard (weak scaling) and the slowest rank counts."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from mcircuit_b200 import circuits
    from mcircuit_b200.engine import B200Engine
    from oracle import cref, restate
    peak = hbm_peak()[0]
    out = []

    # C4's oracle run (the C restatement, sampled lane words, ALL steps) takes about as long as everything else
    # here together: it runs beside the GPU work in a thread (the C call releases the GIL)
    c4_lanes, c4_steps = 1 << 20, a.c4_steps
    c4_root = circuits.lfsr_pipeline()
    c4_seeds = ["/l%d_q%d/out" % (n, b) for n in range(256) for b in (0, 5, 17)]
    c4_words = c4_lanes // 64
    c4_cols = np.array([0, 1, 2, 3, c4_words // 2, c4_words // 2 + 1, c4_words - 2, c4_words - 1], dtype=np.int64)
    c4_sim = cref.CRef(c4_root, 1, True, lanes=64 * 8, threads=1)
    for k, p in enumerate(c4_seeds):
        c4_sim.set_planes(p, restate.stimulus_word(k, c4_cols + rank * c4_words)[None, :])
    c4_oracle = threading.Thread(target=lambda: c4_sim.run(c4_steps), daemon=True)
    c4_oracle.start()

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize(dev)
        best = 1e30
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # prime the stream with ~0.1 ms of spinning so that e0, the launch(es) and e1 are all queued while the GPU is
            # busy: the events then bracket DEVICE time, not the Python -> ctypes -> launch latency of an idle queue
            # (which is what a 20 us single launch would otherwise measure)
            torch.cuda._sleep(200_000)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        if world > 1:
            t = torch.tensor([best], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t[0])
        return best

    def all_ok(ok):
        if world > 1:
            t = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(int(t[0]))
        return bool(ok)

    def port_rate(root, lanes, steps, threads=None):
        sim = cref.CRef(root, 1, True, lanes=lanes, threads=threads)
        sim.run(1)
        t = time.perf_counter()
        sim.run(steps)
        dt = time.perf_counter() - t
        return sim, dt

    def entry(name, eng, lanes, steps, ms, bound, parity, scales, extra=None, executed_steps=None):
        g = eng.program.n_gates
        # bytes are counted for the steps that were EXECUTED: skipped (idle) steps move nothing
        alg = eng.algorithmic_bytes(steps if executed_steps is None else executed_steps) / (ms * 1e-3) / 1e9
        e = {"name": name, "mode": eng.mode, "lanes_per_gpu": lanes, "lanes_total": lanes * (world if scales else 1),
             "steps": steps, "gates": g, "ms": ms, "ms_per_step": ms / steps,
             "value": g * lanes * (world if scales else 1) * steps / (ms * 1e-3), "unit": UNIT,
             "roofline": {"bound": bound, "achieved_GBps_algorithmic": alg, "frac_of_hbm_peak": alg / peak},
             "parity": parity, "scales_with_gpus": scales}
        if extra:
            e.update(extra)
        out.append(e)
        return e

    # ---- C1: examples/counter.py, 1,000 cycles = 2,000 steps, single instance ---------------------------
    root = circuits.counter_circuit()
    eng = B200Engine(root, 501, True, lanes=64, device=dev.index)
    ora = restate.RestatedJIT(root, 501, True)
    eng.run(2000)
    for _ in range(2000):
        ora.step()
    pins = ("/clk/out", "/r/out", "/r/prevclock")
    ok = all(eng.get_pin_state(p) == ora.get_pin_state(p) for p in pins)
    eng.burst()
    ora.burst()
    ok = ok and all(eng.get_pin_state(p) == ora.get_pin_state(p) for p in pins)
    ms = timed(lambda: eng.run(2000))
    e = entry("C1 examples/counter.py clocked counter, 1,000 cycles (2,000 steps), single instance", eng, 64, 2000, ms,
              "latency (one warp walks 4 micro-ops per step; nothing to parallelise in a single instance; mode chosen: %s)" % eng.mode,
              {"bit_exact": all_ok(ok), "against": "oracle/restate.py after 2,000 steps and after burst() (502 more)", "pins": list(pins)},
              False, {"lanes_note": "the reference's single instance is lane 0 of one 64-lane word; replicas only at N > 1"})
    if rank == 0:
        sim, dt = port_rate(root, 64, 200_000, threads=1)
        e["cpu_port"] = {"value": 4 * 64 * 200_000 / dt, "cores": 1, "kind": "port", "sample": "64 lanes, 200,000 steps"}
        e["cpu_jit"] = jits.get("c1")
    del eng

    # ---- C2: 32-bit ripple-carry adder, 2^20 vectors, one combinational step -----------------------------
    lanes = 1 << 20
    root = circuits.ripple_adder(nbits=32)
    eng = B200Engine(root, 1, True, lanes=lanes, device=dev.index, keep="ports", lane_word_offset=rank * (lanes // 64))
    eng.randomize_inputs()
    eng.step()
    planes = {p: eng.get_pin_planes(p)[0] for p in eng.input_pins() + eng.output_pins()}
    carry = planes["/cin/pin"].copy()
    ok = True
    for k in range(32):                                   # a + b + cin, bit-sliced, on ALL 2^20 lanes
        x, y = planes["/a%d/pin" % k], planes["/b%d/pin" % k]
        ok = ok and np.array_equal(planes["/s%d/pin" % k], x ^ y ^ carry)
        carry = (x & y) | (carry & (x ^ y))
    ok = ok and np.array_equal(planes["/cout/pin"], carry)
    ms = timed(lambda: eng.run(1), reps=5)
    io_bytes = (65 + 33) * (lanes // 64) * 8
    e = entry("C2 32-bit ripple-carry adder (160 gates), 2^20 random vectors, one step", eng, lanes, 1, ms,
              "latency + compulsory I/O (65 dependent micro-op levels; %.1f MB of operand and result rows)" % (io_bytes / 1e6),
              {"bit_exact": all_ok(ok), "against": "a + b + cin recomputed bit-sliced in numpy on all 2^20 lanes (sum and true carry-out)"},
              True, {"compulsory_io_GBps": io_bytes / (ms * 1e-3) / 1e9})
    if rank == 0:
        sim, dt = port_rate(root, 1 << 20, 4)
        e["cpu_port"] = {"value": 160 * (1 << 20) * 4 / dt, "cores": sim.threads, "kind": "port", "sample": "2^20 lanes, 4 passes"}
        e["cpu_jit"] = [jits.get("c2_w1"), jits.get("c2_w64")]
    del eng
    torch.cuda.empty_cache()

    # ---- C3: 64x64 array multiplier, 2^24 vectors, one step -------------------------------------------------
    lanes = 1 << 24
    root = circuits.array_multiplier(n=64)
    eng = B200Engine(root, 1, True, lanes=lanes, device=dev.index, mode="levels", keep="ports", lane_word_offset=rank * (lanes // 64))
    eng.randomize_inputs()
    eng.step()
    words = lanes // 64
    cols = np.array([0, 1, words // 2, words - 1], dtype=np.int64)

    def sample(pins):
        return eng.sample_planes(pins, cols)
    ia, ib = ["/a%d/pin" % k for k in range(64)], ["/b%d/pin" % k for k in range(64)]
    po = ["/p%d/pin" % k for k in range(128)]
    va, vb, vp = restate.unpack_lanes(sample(ia)), restate.unpack_lanes(sample(ib)), sample(po)
    plo, phi = restate.unpack_lanes(vp[:64]), restate.unpack_lanes(vp[64:])
    ok = all((int(x) * int(y)) == (int(h) << 64 | int(l)) for x, y, l, h in zip(va, vb, plo, phi))
    ms = timed(lambda: eng.run(1))
    e = entry("C3 64x64 array multiplier (24,064 gates), 2^24 random vectors bit-sliced, one step", eng, lanes, 1, ms,
              "hbm (level kernel; a level is ~64 gates, the previous level's rows are still in L2)",
              {"bit_exact": all_ok(ok), "against": "a * b in Python integers on %d sampled lanes (4 lane words of this shard)" % len(va)}, True)
    if rank == 0:
        sim, dt = port_rate(root, 1 << 16, 2)
        e["cpu_port"] = {"value": eng.program.n_gates * (1 << 16) * 2 / dt, "cores": sim.threads, "kind": "port", "sample": "2^16 lanes, 2 passes"}
        e["cpu_jit"] = {"unavailable": "the hierarchical 64x64 build takes 138 s of reference JIT build time (see tests/golden: its result is pinned there); not repeated per bench run"}
    del eng
    torch.cuda.empty_cache()

    # ---- C4: 256 LFSRs + 1,000-register pipeline, state resident on the device ------------------------------
    lanes, steps, root, seeds, words, cols = c4_lanes, c4_steps, c4_root, c4_seeds, c4_words, c4_cols
    eng = B200Engine(root, 1, True, lanes=lanes, device=dev.index, mode="serial", keep="ports", share_edge_detectors=True,
                     lane_word_offset=rank * (lanes // 64))
    eng.randomize_inputs(seeds)
    eng.rt.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    eng.run(steps)                                        # ONE launch, no host round trip
    e1.record()
    torch.cuda.synchronize(dev)
    launches = eng.rt.launch_count()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    c4_oracle.join()
    sim = c4_sim
    probes = ["/pipe_out/pin"] + ["/tap%d/pin" % n for n in range(0, 256, 15)]
    got = eng.sample_planes(probes, cols)
    ok = all(np.array_equal(got[k], sim.planes(p)[0]) for k, p in enumerate(probes))
    ones = float(np.unpackbits(got.view(np.uint8)).mean())       # the end state is alive, not a comparison of zeros
    ok = ok and 0.4 < ones < 0.6
    full = steps == 200_000
    e = entry("C4 256 LFSRs + 1,000-register pipeline, %s clock cycles (%d steps)%s, 2^20 lanes, state resident"
              % ("{:,}".format(steps // 2), steps, " = the named size" if full else " of the 10^5 named"), eng, lanes, steps, ms,
              "latency / issue (4 consumer warps per SM walk the whole stream; at 2^20 lanes there are only 512 warps of work); "
              "rising steps read and write every register row once (TMA-staged), idle steps are skipped (event-driven: the "
              "clock's falling phase is the identity on every register)",
              {"bit_exact": all_ok(ok), "against": "oracle/mcref.c run for the same %d steps on 8 sampled lane words (512 lanes), "
                                                   "pipe_out and 18 LFSR taps" % steps,
               "ones_fraction_of_the_compared_words": ones,
               "register_note": "Register parity is unpinned by the reference (its emitter does not compile, SURVEY Q3)"}, True,
              {"launches": int(launches), "regions_skipped_by_cta0": eng.regions_skipped(),
               "micro_ops_note": "gates counts the micro-ops executed with shared clock edge detectors (10,707); the unshared "
                                 "lowering has 29,089", "serial": eng.program.serial_info,
               "full_size": full,
               "executed_steps": steps - eng.regions_skipped(),
               "roofline_note": "value counts every gate of every step (the metric's definition: an idle clock phase is evaluated "
                                "exactly, by skipping it); roofline.achieved counts bytes of the EXECUTED (rising) steps only"},
              executed_steps=steps - eng.regions_skipped())
    if rank == 0:
        sim2, dt = port_rate(root, 1 << 16, 20)
        e["cpu_port"] = {"value": 29089 * (1 << 16) * 20 / dt, "cores": sim2.threads, "kind": "port",
                         "sample": "2^16 lanes, 20 steps, 29,089 micro-ops per step (no sharing in the port)",
                         "steps_per_s_at_2^16_lanes": 20 / dt}
        e["cpu_jit"] = {"unavailable": "Register does not compile in the reference (core/simulator.py:121-126, SURVEY Q3)"}
    del eng
    torch.cuda.empty_cache()
    return out


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
    jit = None
    if rank == 0 and not a.no_jit and not a.no_cpu_baseline:
        cases = JIT_CASES_C5 + (JIT_CASES_SMALL if not a.no_configs else ())
        jit = JitWorkers(cases, a.jit_seconds)            # they build while the GPU works, then wait
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

    # ---- parity of this very run (outside the timed region), every rank's shard -------------------------
    verify, port_sim = None, None
    if not a.no_verify:
        if rank == 0 and world == 1 and not a.no_cpu_baseline:
            port_sim = make_port(a, root)                # one oracle instance serves the check and the CPU baseline
        verify, _ = verify_c5(a, eng, rank, words, port_sim)
        if world > 1:
            t = torch.tensor([1 if verify["bit_exact"] else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            verify["bit_exact_all_ranks"] = bool(int(t[0]))
            verify["ranks_checked"] = world
        else:
            verify["bit_exact_all_ranks"] = verify["bit_exact"]
            verify["ranks_checked"] = 1

    # ---- roofline of the dominant kernel (k_eval_level), measured live ------------------------------
    peak, peak_src = hbm_peak()
    level_launches = prog.n_levels * eng.n_tiles            # per step per GPU
    alg_bytes_step = eng.algorithmic_bytes(1)               # this GPU, one step
    agg_launch_s = (eval_total * 1e-3) / (a.steps * level_launches)
    achieved = (alg_bytes_step / level_launches) / agg_launch_s / 1e9
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tp):
        try:
            traffic = json.load(open(tp))
        except Exception:
            traffic = {}
    alg_launch = alg_bytes_step / level_launches
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("steady_dram_bytes_per_launch", traffic.get("k_eval_level_dram_bytes_per_launch")),
                "traffic_cold": traffic.get("k_eval_level_dram_bytes_per_launch"),
                "traffic_steady": traffic.get("steady_dram_bytes_per_launch"),
                "traffic_source": traffic.get("steady_source", traffic.get("source")),
                "kernel": "k_eval_level", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_launch,
                "aggregate_us_per_launch": agg_launch_s * 1e6, "launches_per_step": level_launches,
                "isolated_kernel_us": traffic.get("isolated_kernel_us"),
                "eval_share_of_step": eval_total / ms,
                "l2_ceiling_GBps": traffic.get("k1_all_l2_hits_GBps"),
                "note": "achieved counts ALGORITHMIC bytes (24 B per gate-word, SURVEY 8d) over the AGGREGATE launch rate (two "
                        "tiles overlap on two streams, so aggregate_us_per_launch is not one kernel's duration - see "
                        "isolated_kernel_us); with %d-word tiles part of the bytes is served by L2 (previous level's rows, "
                        "recycled scratch slots), so frac exceeds 1 while real DRAM traffic (traffic_steady) stays below the "
                        "peak.  l2_ceiling_GBps is the same kernel with every fan-in row an L2 hit "
                        "(profiles/r02_probe_k1_l2_ceiling.jsonl): the bound no tiling can pass" % eng.tile_words}

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
        h2d, d2h = int(h_in.numel() * 8), int((h_out.numel() + h_sig.numel()) * 8)
        e2e = {"value": prog.n_gates * total_lanes * a.steps / (ems * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ems / a.steps,
               "host_link_GBps_per_gpu": (h2d + d2h) / (ems / a.steps * 1e-3) / 1e9,
               "host_link_GBps_all_gpus": world * (h2d + d2h) / (ems / a.steps * 1e-3) / 1e9,
               "host_link_ceiling": (traffic.get("host_copy_ceiling") or {}).get(str(world)),
               "host_link_ceiling_note": (traffic.get("host_copy_ceiling") or {}).get("what"),
               "api": "B200Engine.run_host(h_in, h_out) [tile-pipelined H2D / eval / D2H over the C ABI] "
                      "-> signature -> D2H (pinned host buffers)"}
        ceil = e2e["host_link_ceiling"]
        if ceil:
            # both directions are in flight at once, so the platform's concurrent figure is the bound
            e2e["host_link_fraction_of_ceiling"] = e2e["host_link_GBps_all_gpus"] / ceil["both_GBps"]
        assert np.array_equal(h_sig.numpy().view(np.uint64), signature_host), "e2e signature differs from resident run"
        # the downloaded planes really are the outputs: their popcounts are the signature
        local_sig = eng.rt.to_host_u64(eng.signature_run(sig))       # this rank's shard, not reduced
        idx = list(range(0, n_out, max(1, n_out // 16)))
        for i in idx:
            pc = int(np.unpackbits(h_out[i].numpy().view(np.uint8)).sum())
            assert pc == int(local_sig[i]), "downloaded output plane %d does not match its signature" % i
        e2e["downloaded_planes_checked_against_signature"] = len(idx)
        del h_in, h_out

    c5_meta = {"levels": prog.n_levels, "tile_words": eng.tile_words, "tiles": eng.n_tiles,
               "persistent_slots": prog.n_persist, "scratch_slots": prog.n_scratch, "mode": eng.mode, "streams": eng.n_streams}
    n_gates = prog.n_gates
    sig_sum = int(signature_host[:sig["n"]].sum())
    eng.close()
    del eng, buf
    torch.cuda.empty_cache()

    # ---- the reference's JIT (released now: the GPU-side host work is over) and configs C1-C4 -----------
    jits = {}
    if jit is not None:
        jit.release()
        jits = jit.collect()
    configs = None
    if not a.no_configs:
        configs = run_configs(a, dev, world, rank, jits)

    cb = None
    if rank == 0 and not a.no_cpu_baseline:
        if world == 1:
            if port_sim is None:
                port_sim = make_port(a, root)
            port_stimulus(port_sim, a, range(port_sim.words))
            cb, _ = time_port(port_sim, a, a.cpu_seconds)
        else:
            cb = {"kind": "port", "unavailable": "timed at N=1 only (the ranks' own host work would disturb it)"}
        cb["port"] = dict(cb)
        cb["jit"] = [jits[c] for c in JIT_CASES_C5 if c in jits]
        cb["jit_note"] = jit_note()
        cb["host_cpus"] = os.cpu_count()

    ok = verify is None or verify["bit_exact_all_ranks"]
    if configs:
        ok = ok and all(c["parity"]["bit_exact"] for c in configs)
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": dict({"workload": workload_name(a), "gates": n_gates, "depth": a.depth,
                            "inputs": a.inputs, "lanes_per_gpu": a.lanes_per_gpu, "lanes_total": total_lanes,
                            "parallelism": "lanes sharded x%d, netlist replicated" % world,
                            "l2": "inputs larger than L2: %.1f GB of state streamed per step" % (alg_bytes_step / 1e9),
                            "stimulus": "splitmix64 counter-based, generated on device",
                            "compile_s": round(t_build, 1), "tuning": a.tune, "numa_node": numa}, **c5_meta),
            "verify": verify, "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "signature_popcount_sum": sig_sum, "configs": configs, "build": build_info(),
        }
        print(json.dumps(out))
    if sampler:
        sampler.stop()
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.exit(3)                                       # a fast result that differs from the reference's is not a result


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
