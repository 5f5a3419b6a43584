"""Multi-GPU: stimulus lanes shard across ranks (one process per GPU), the netlist is
replicated, and nothing is exchanged during evaluation - lanes never interact.
The only collective is one all-reduce SUM over the per-rank signature / toggle vectors
(NCCL over NVLink on GPUs; gloo in the CPU tests).  Sums are taken on the int64 view:
two's-complement wrap-around is exactly uint64 addition mod 2^64.
"""
from __future__ import annotations

import numpy as np


def shard_lanes(total_lanes: int, rank: int, world: int):
    """(lanes, first lane word) of ``rank``: contiguous shards of whole 64-lane words, the
    remainder spread over the first ranks."""
    if total_lanes % 64:
        raise ValueError("total_lanes must be a multiple of 64")
    words = total_lanes // 64
    base, extra = divmod(words, world)
    mine = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return mine * 64, first


def all_reduce_sum(buf, group=None):
    """In-place SUM of a uint64 vector held as an int64 torch tensor or a numpy array."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return buf
    if isinstance(buf, np.ndarray):
        t = torch.from_numpy(buf.view(np.int64))          # shares memory
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return buf
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


class ShardedEngine:
    """The rank-local engine of a lane-sharded run + the reduced observables."""

    def __init__(self, root, total_lanes, burst_size=500, map_pins=True, group=None, **engine_kw):
        import torch.distributed as dist
        from .engine import B200Engine
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        lanes, first_word = shard_lanes(total_lanes, self.rank, self.world)
        if lanes == 0:
            raise ValueError("fewer lane words than ranks")
        self.total_lanes = total_lanes
        self.engine = B200Engine(root, burst_size, map_pins, lanes=lanes, lane_word_offset=first_word,
                                 **engine_kw)
        self._sig_plans = {}

    def randomize_inputs(self, *a, **kw):
        return self.engine.randomize_inputs(*a, **kw)

    def run(self, steps):
        self.engine.run(steps)

    def signature(self, pins=None):
        """(popcount, checksum) over ALL lanes of all ranks."""
        key = tuple(pins) if pins is not None else None
        plan = self._sig_plans.get(key)
        if plan is None:
            plan = self._sig_plans[key] = self.engine.signature_plan(pins)
        buf = all_reduce_sum(self.engine.signature_run(plan), self.group)
        host = self.engine.rt.to_host_u64(buf)
        n, half = plan["n"], plan["half"]
        return host[:n], host[half:half + n]

    def toggle_counts(self):
        buf = all_reduce_sum(self.engine.toggle_counts_device().clone()
                             if hasattr(self.engine.toggle_counts_device(), "clone")
                             else self.engine.toggle_counts_device().copy(), self.group)
        return self.engine.rt.to_host_u64(buf)
