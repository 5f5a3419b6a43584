"""not-gpu: oracle/restate.py and our flattening helpers against the LIVE reference
(/root/reference through oracle/refshim.py).  Skipped where the reference is absent
(the GPU box); tests/golden/ pins the same behaviour there."""
import random

import numpy as np
import pytest

import mcircuit_b200.core.descriptors as MD
from mcircuit_b200 import circuits
from oracle import refshim, restate


def _strip(topology):
    return [(k, (v[1] if k == "emit" else v)) for k, v in topology]


@pytest.mark.parametrize("seed", range(12))
def test_schedule_pins_connections_sources_match_reference(reference, seed):
    D, S = reference
    root, _ = circuits.random_circuit(D, seed, n_nodes=16)
    assert list(S.iter_simulation_pins(root)) == list(restate._pins(root))
    assert list(S.iter_simulation_connections(root)) == list(restate._connections(root))
    assert _strip(S.iter_simulation_topology(root)) == _strip(restate._topology(root))
    assert S._map_to_sources(root) == restate._sources(root)


@pytest.mark.parametrize("seed", range(12))
def test_our_core_simulator_helpers_match_reference(reference, seed):
    """mcircuit_b200.core.simulator re-exports the reference's module-level helpers; same
    circuit built with each descriptor module must give identical listings."""
    D, S = reference
    from mcircuit_b200.compiler import flatten
    ref_root, _ = circuits.random_circuit(D, seed, n_nodes=16)
    own_root, _ = circuits.random_circuit(MD, seed, n_nodes=16)
    d = flatten(own_root, True)
    assert list(S.iter_simulation_pins(ref_root)) == list(d.iter_pins())
    conns = [(d.pin_path(s), d.pin_path(t)) for s, t in zip(d.conn_src, d.conn_dst)]
    assert sorted(S.iter_simulation_connections(ref_root)) == sorted(conns)
    emits = [d.leaf_path[a] for k, a in d.schedule if k == 0]
    assert [v[1] for k, v in S.iter_simulation_topology(ref_root) if k == "emit"] == emits
    want = S._map_to_sources(ref_root)
    got = {d.pin_path(p): (None if d.pin_src[p] < 0 else d.pin_path(int(d.pin_src[p])))
           for p in range(len(d.pin_width))}
    assert want == got


@pytest.mark.parametrize("seed", range(40, 60))
@pytest.mark.parametrize("map_pins", [True, False])
def test_oracle_equals_live_jit_on_random_circuits(reference, seed, map_pins):
    D, _ = reference
    root, probes = circuits.random_circuit(D, seed, n_nodes=20)
    jit = refshim.build_jit(root, 5, map_pins)
    ora = restate.RestatedJIT(root, 5, map_pins)
    rng = random.Random(seed)
    for step in range(12):
        if step == 5:
            for p in ("/in0/pin", "/in1/pin"):
                v = rng.getrandbits(ora._width[p])
                jit.set_pin_state(p, v)
                ora.set_pin_state(p, v)
        jit.step()
        ora.step()
        for p in probes:
            assert jit.get_pin_state(p) == ora.get_pin_state(p), (step, p)
    jit.burst()
    ora.burst()
    for p in probes:
        assert jit.get_pin_state(p) == ora.get_pin_state(p), ("burst", p)


def test_burst_is_burst_size_plus_one(reference):
    D, _ = reference
    root = circuits.counter_circuit(D)
    jit = refshim.build_jit(root, 501, True)
    ora = restate.RestatedJIT(root, 501, True)
    for _ in range(8):
        jit.step()
    jit.burst()          # 502 more steps: period 4 -> net +2
    for _ in range(8 + 502):
        ora.step()
    for p in ("/clk/out", "/r/out", "/r/prevclock"):
        assert jit.get_pin_state(p) == ora.get_pin_state(p)


def test_lane_trick_c2_adder_is_real_addition(reference):
    """The reference at width=64 on the 160-gate adder computes a+b+cin for 64 packed lanes."""
    D, _ = reference
    nb = 16
    jit = refshim.build_jit(circuits.ripple_adder(D, nb, width=64), 1, True)
    rng = np.random.default_rng(3)
    a = rng.integers(0, 1 << nb, size=64, dtype=np.uint64)
    b = rng.integers(0, 1 << nb, size=64, dtype=np.uint64)
    cin = rng.integers(0, 2, size=64, dtype=np.uint64)
    pa, pb, pc = restate.pack_lanes(a, nb), restate.pack_lanes(b, nb), restate.pack_lanes(cin, 1)
    jit.set_pin_state("/cin/pin", int(pc[0, 0]))
    for i in range(nb):
        jit.set_pin_state(f"/a{i}/pin", int(pa[i, 0]))
        jit.set_pin_state(f"/b{i}/pin", int(pb[i, 0]))
    jit.step()
    planes = np.array([[jit.get_pin_state(f"/s{i}/pin")] for i in range(nb)] + [[jit.get_pin_state("/cout/pin")]],
                      dtype=np.uint64)
    total = restate.unpack_lanes(planes)
    assert np.array_equal(total, a + b + cin)
