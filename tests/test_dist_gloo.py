"""not-gpu: the N>1 path on CPU - two processes, gloo, lanes sharded, signatures reduced.
The kernels are replaced by the numpy test double; what is under test is the host logic:
shard geometry, lane_word_offset in stimulus and checksums, and the all-reduce."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_lanes, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from _cpu_runtime import CpuRuntime
    from mcircuit_b200 import circuits
    from mcircuit_b200.dist import ShardedEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        root = circuits.random_dag(n_gates=600, depth=12, n_inputs=24, seed=4)
        sh = ShardedEngine(root, total_lanes, 1, True, runtime=CpuRuntime(), keep="ports", mode="levels")
        sh.randomize_inputs()
        sh.run(1)
        pc, cs = sh.signature()
        clk = ShardedEngine(circuits.counter_circuit(), total_lanes, 1, True, runtime=CpuRuntime())
        clk.engine.enable_toggle_counts(["/clk/out"])
        clk.run(6)
        tg = clk.toggle_counts()
        # C4's shape through the stream kernel's program (mode='serial'): per-rank seeds come from the counter-based
        # stimulus at the rank's GLOBAL lane words, signatures are reduced over the ranks
        seq = circuits.lfsr_pipeline(n_lfsr=4, lfsr_bits=8, stages=24, taps=(7, 5, 1, 0))
        ser = ShardedEngine(seq, total_lanes, 1, True, runtime=CpuRuntime(), keep="ports", mode="serial")
        ser.randomize_inputs(["/l%d_q%d/out" % (n, b) for n in range(4) for b in (0, 3)])
        ser.run(21)
        spc, scs = ser.signature()
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), pc=pc, cs=cs, tg=tg, spc=spc, scs=scs,
                 smode=ser.engine.mode, lanes=sh.engine.lanes, off=sh.engine.lane_word_offset)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total_lanes", [(2, 64 * 10), (2, 64 * 7)])
def test_two_rank_sharded_signature_equals_single_engine(tmp_path, world, total_lanes):
    from _cpu_runtime import CpuRuntime
    from mcircuit_b200 import circuits
    from mcircuit_b200.engine import B200Engine
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total_lanes, str(tmp_path)), nprocs=world, join=True)
    root = circuits.random_dag(n_gates=600, depth=12, n_inputs=24, seed=4)
    one = B200Engine(root, 1, True, lanes=total_lanes, runtime=CpuRuntime(), keep="ports", mode="levels")
    one.randomize_inputs()
    one.step()
    pc, cs = one.signature()
    shards = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert sum(int(s["lanes"]) for s in shards) == total_lanes
    assert [int(s["off"]) for s in shards] == [0, int(shards[0]["lanes"]) // 64]
    seq = circuits.lfsr_pipeline(n_lfsr=4, lfsr_bits=8, stages=24, taps=(7, 5, 1, 0))
    ser = B200Engine(seq, 1, True, lanes=total_lanes, runtime=CpuRuntime(), keep="ports", mode="serial")
    ser.randomize_inputs(["/l%d_q%d/out" % (n, b) for n in range(4) for b in (0, 3)])
    ser.run(21)
    spc, scs = ser.signature()
    assert spc.sum() > 0
    for s in shards:                       # every rank holds the reduced (global) vectors
        assert np.array_equal(s["pc"], pc) and np.array_equal(s["cs"], cs)
        assert int(s["tg"][0]) == 6 * total_lanes
        assert str(s["smode"]) == "serial" and np.array_equal(s["spc"], spc) and np.array_equal(s["scs"], scs)


def test_shard_lanes_geometry():
    from mcircuit_b200.dist import shard_lanes
    assert shard_lanes(64 * 8, 0, 8) == (64, 0) and shard_lanes(64 * 8, 7, 8) == (64, 7)
    got = [shard_lanes(64 * 10, r, 4) for r in range(4)]
    assert got == [(192, 0), (192, 3), (128, 6), (128, 8)]
    assert sum(g[0] for g in got) == 640
    with pytest.raises(ValueError):
        shard_lanes(100, 0, 2)
