"""Platform ceiling of pinned host <-> device copies: every rank (one per GPU) moves 64 MB blocks
with plain 1-D cudaMemcpyAsync (torch copy_ on pinned tensors), H2D alone, D2H alone and both at
once on two streams; all ranks concurrently.  Prints per-rank and aggregate GB/s.
  python -m torch.distributed.run --nproc-per-node N scripts/r02_probe_hostcopy.py   (or plain python for N=1)"""
import json, os, sys, time
import torch
import torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29511")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
BLOCK = 64 << 20
NB = 16                                   # 1 GiB each way per repetition
h_in = torch.empty((NB, BLOCK // 8), dtype=torch.int64, pin_memory=True); h_in.fill_(1)
h_out = torch.empty((NB, BLOCK // 8), dtype=torch.int64, pin_memory=True); h_out.fill_(2)
d_in = torch.empty((NB, BLOCK // 8), dtype=torch.int64, device=dev)
d_out = torch.ones((NB, BLOCK // 8), dtype=torch.int64, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

def sync():
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()

def run(h2d, d2h, reps=4):
    sync()
    t = time.perf_counter()
    for _ in range(reps):
        for b in range(NB):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[b].copy_(h_in[b], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[b].copy_(d_out[b], non_blocking=True)
    torch.cuda.synchronize(dev)
    dt_local = time.perf_counter() - t
    sync()
    dt = time.perf_counter() - t
    moved = reps * NB * BLOCK * (int(h2d) + int(d2h))
    return moved / dt_local / 1e9, moved * world / dt / 1e9

run(True, True, 1)
res = {}
for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    res[name] = run(a, b)
if world > 1:
    t = torch.tensor([res[k][0] for k in ("h2d", "d2h", "both")], device=dev, dtype=torch.float64)
    lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    per_rank_min = lo.tolist()
else:
    per_rank_min = [res[k][0] for k in ("h2d", "d2h", "both")]
if rank == 0:
    print(json.dumps({"probe": "pinned_host_copy_ceiling", "ranks": world, "block_MB": BLOCK >> 20,
                      "aggregate_GBps": {k: res[k][1] for k in res}, "slowest_rank_GBps": dict(zip(("h2d", "d2h", "both"), per_rank_min)),
                      "cpus": os.cpu_count()}), flush=True)
if world > 1:
    dist.destroy_process_group()
