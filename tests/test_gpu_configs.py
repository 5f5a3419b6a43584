"""-m gpu: the five BASELINE.json configs.  Small cases are compared in full with the oracle;
full-size cases through size-independent properties (the adder adds, the multiplier
multiplies, tiles agree with each other, signatures are sums) plus sampled columns against
the plain-C oracle."""
import numpy as np
import pytest

import mcircuit_b200.core.descriptors as MD
from mcircuit_b200 import circuits
from mcircuit_b200.engine import B200Engine
from oracle import cref, restate

pytestmark = pytest.mark.gpu
U64 = np.uint64


def planes_to_int(eng, fmt, n):
    """per-lane integer (object array) from n one-bit ports"""
    total = np.zeros(eng.lanes, dtype=object)
    for i in range(n):
        total += eng.get_pin_state(fmt % i, lanes=True).astype(object) << i
    return total


# ---- C1: examples/counter.py, 1,000 cycles, single instance -------------------------------------
def test_c1_counter_1000_cycles():
    eng = B200Engine(circuits.counter_circuit(MD), 501, True)          # lanes=64, lane 0 = the instance
    ora = restate.RestatedJIT(circuits.counter_circuit(MD), 501, True)
    for step in range(2000):                                           # 1,000 clock cycles
        eng.step()
        ora.step()
        if step % 97 == 0 or step > 1990:
            for p in ("/clk/out", "/r/out", "/r/prevclock"):
                assert eng.get_pin_state(p) == ora.get_pin_state(p), (step, p)
    eng.burst()
    ora.burst()
    for p in ("/clk/out", "/r/out", "/r/prevclock"):
        assert eng.get_pin_state(p) == ora.get_pin_state(p)
    # the same through one resident launch
    eng2 = B200Engine(circuits.counter_circuit(MD), 501, True)
    eng2.run(2000 + 502)
    for p in ("/clk/out", "/r/out", "/r/prevclock"):
        assert eng2.get_pin_state(p) == ora.get_pin_state(p)


# ---- C2: 32-bit ripple-carry adder, 2^20 random vectors --------------------------------------------
@pytest.mark.parametrize("mode", ["resident", "levels"])
def test_c2_adder_2_20_vectors(mode):
    lanes = 1 << 20
    eng = B200Engine(circuits.ripple_adder(MD, 32), 1, True, lanes=lanes, mode=mode)
    rng = np.random.default_rng(20)
    a = rng.integers(0, 1 << 32, size=lanes, dtype=U64)
    b = rng.integers(0, 1 << 32, size=lanes, dtype=U64)
    cin = rng.integers(0, 2, size=lanes, dtype=U64)
    pa, pb = restate.pack_lanes(a, 32), restate.pack_lanes(b, 32)
    eng.set_pin_state("/cin/pin", cin)
    for i in range(32):
        eng.set_pin_planes(f"/a{i}/pin", pa[i][None, :])
        eng.set_pin_planes(f"/b{i}/pin", pb[i][None, :])
    eng.step()
    planes = np.stack([eng.get_pin_planes(f"/s{i}/pin")[0] for i in range(32)] + [eng.get_pin_planes("/cout/pin")[0]])
    assert np.array_equal(restate.unpack_lanes(planes), a + b + cin)
    pc, _ = eng.signature()
    want = np.array([np.unpackbits(row.view(np.uint8)).sum() for row in planes], dtype=U64)
    assert np.array_equal(pc, want)


# ---- C3: 64x64 array multiplier (~30k gates) ------------------------------------------------------------
def test_c3_multiplier_small_vs_oracle():
    root = circuits.array_multiplier(MD, 8)
    lanes = 64 * 8
    eng = B200Engine(root, 1, True, lanes=lanes, keep="all", mode="levels")
    ora = cref.CRef(root, 1, True, lanes=lanes)
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, size=lanes, dtype=U64)
    b = rng.integers(0, 256, size=lanes, dtype=U64)
    for i in range(8):
        for name, v in (("a", a), ("b", b)):
            bit = (v >> U64(i)) & U64(1)
            eng.set_pin_state(f"/{name}{i}/pin", bit)
            ora.set_pin_state(f"/{name}{i}/pin", bit)
    eng.step()
    ora.step()
    for k in range(16):
        assert np.array_equal(eng.get_pin_planes(f"/p{k}/pin"), ora.planes(f"/p{k}/pin")), k
    assert np.array_equal(planes_to_int(eng, "/p%d/pin", 16), a.astype(object) * b.astype(object))


def test_c3_multiplier_64x64_multiplies():
    root = circuits.array_multiplier(MD, 64)
    lanes = 1 << 16                     # full netlist; lanes bounded so the check stays in seconds
    eng = B200Engine(root, 1, True, lanes=lanes, keep="ports", mode="levels")
    assert 20_000 < eng.program.n_gates < 32_000
    eng.randomize_inputs()
    eng.step()
    a = planes_to_int(eng, "/a%d/pin", 64)
    b = planes_to_int(eng, "/b%d/pin", 64)
    p = planes_to_int(eng, "/p%d/pin", 128)
    assert (p == a * b).all()


# ---- C4: LFSRs + register pipeline, state resident across cycles ----------------------------------------------
@pytest.mark.parametrize("mode", ["resident", "levels"])
def test_c4_lfsr_pipeline_vs_oracle(mode):
    args = dict(n_lfsr=8, lfsr_bits=16, stages=40, taps=(15, 13, 12, 10))
    root = circuits.lfsr_pipeline(MD, **args)
    lanes = 256
    eng = B200Engine(root, 1, True, lanes=lanes, mode=mode)
    ora = cref.CRef(root, 1, True, lanes=lanes)
    rng = np.random.default_rng(4)
    for n in range(8):
        for b in range(16):
            v = rng.integers(0, 2, size=lanes, dtype=U64)
            eng.set_pin_state(f"/l{n}_q{b}/out", v)
            ora.set_pin_state(f"/l{n}_q{b}/out", v)
    probes = [f"/tap{n}/pin" for n in range(8)] + ["/pipe_out/pin", "/l3_q7/out", "/p17/out", "/p17/prevclock"]
    for chunk in (1, 1, 2, 7, 100, 333):
        eng.run(chunk)
        ora.run(chunk)
        for p in probes:
            assert np.array_equal(eng.get_pin_planes(p), ora.planes(p)), (chunk, p)


def test_c4_full_shape_resident_run():
    """256 LFSRs x 32 + 1,000-register pipeline; state stays on the device for the whole run
    (no host round trip between cycles); lanes and cycles bounded to keep the test short."""
    root = circuits.lfsr_pipeline(MD)
    lanes = 64 * 16
    eng = B200Engine(root, 1, True, lanes=lanes, mode="resident", keep="ports")
    ora = cref.CRef(root, 1, True, lanes=lanes)
    rng = np.random.default_rng(5)
    for n in range(0, 256, 5):
        v = rng.integers(0, 2, size=lanes, dtype=U64)
        eng.set_pin_state(f"/l{n}_q0/out", v)
        ora.set_pin_state(f"/l{n}_q0/out", v)
    eng.rt.launch_count(reset=True)
    eng.run(2 * 50)                       # 50 clock cycles, one launch
    assert eng.rt.launch_count() == 1
    ora.run(2 * 50)
    for p in ["/pipe_out/pin"] + [f"/tap{n}/pin" for n in range(0, 256, 17)]:
        assert np.array_equal(eng.get_pin_planes(p), ora.planes(p)), p


# ---- C5: random levelized DAG ---------------------------------------------------------------------------------
@pytest.mark.parametrize("tile_words", [None, 64])
def test_c5_random_dag_20k_vs_oracle(tile_words):
    root = circuits.random_dag(MD, n_gates=20_000, depth=40, n_inputs=256, seed=77)
    lanes = 64 * 320
    eng = B200Engine(root, 1, True, lanes=lanes, keep="ports", mode="levels", tile_words=tile_words)
    eng.randomize_inputs()
    eng.step()
    ora = cref.CRef(root, 1, True, lanes=lanes)
    ora.set_rows(ora._base("/i0/pin"), restate.stimulus_word(np.arange(256)[:, None], np.arange(eng.words)[None, :]))
    ora.step()
    first = ora._base("/o0/pin")
    want = ora.rows(first, 500)
    got = np.stack([eng.get_pin_planes(p)[0] for p in eng.output_pins()])
    assert np.array_equal(got, want)
    pc, cs = eng.signature()
    assert np.array_equal(pc, np.array([np.unpackbits(r.view(np.uint8)).sum() for r in want], dtype=U64))


def test_c5_signature_is_a_sum_over_shards():
    """Two half-width engines with lane_word_offset == one full-width engine (what the
    multi-GPU all-reduce relies on)."""
    root = circuits.random_dag(MD, n_gates=4000, depth=20, n_inputs=64, seed=8)
    full = B200Engine(root, 1, True, lanes=64 * 96, keep="ports", mode="levels")
    full.randomize_inputs()
    full.step()
    pc, cs = full.signature()
    acc_pc, acc_cs = np.zeros_like(pc), np.zeros_like(cs)
    for r in range(2):
        part = B200Engine(root, 1, True, lanes=64 * 48, keep="ports", mode="levels", lane_word_offset=48 * r)
        part.randomize_inputs()
        part.step()
        p, c = part.signature()
        with np.errstate(over="ignore"):
            acc_pc += p
            acc_cs += c
    assert np.array_equal(pc, acc_pc) and np.array_equal(cs, acc_cs)


@pytest.mark.parametrize("mode", ["resident", "levels"])
def test_c4_shared_edge_detectors_vs_oracle(mode):
    args = dict(n_lfsr=8, lfsr_bits=16, stages=40, taps=(15, 13, 12, 10))
    root = circuits.lfsr_pipeline(MD, **args)
    lanes = 256
    eng = B200Engine(root, 1, True, lanes=lanes, mode=mode, share_edge_detectors=True)
    ora = cref.CRef(root, 1, True, lanes=lanes)
    rng = np.random.default_rng(4)
    for n in range(8):
        for b in range(16):
            v = rng.integers(0, 2, size=lanes, dtype=U64)
            eng.set_pin_state(f"/l{n}_q{b}/out", v)
            ora.set_pin_state(f"/l{n}_q{b}/out", v)
    probes = [f"/tap{n}/pin" for n in range(8)] + ["/pipe_out/pin", "/l3_q7/out", "/p17/out", "/p17/prevclock",
                                                    "/l0_q0/prevclock", "/l7_q15/prevclock"]
    for chunk in (1, 1, 2, 7, 100):
        eng.run(chunk)
        ora.run(chunk)
        for p in probes:
            assert np.array_equal(eng.get_pin_planes(p), ora.planes(p)), (chunk, p)
