mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/r02_topo_n${N}.txt 2>&1; free -g | head -2 >> gpurun_out/r02_topo_n${N}.txt; nproc >> gpurun_out/r02_topo_n${N}.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/r02_probe_hostcopy.py > gpurun_out/r02_probe_hostcopy_n${N}.jsonl 2> gpurun_out/r02_probe_hostcopy_n${N}.err; echo "probe rc=$?"
cat gpurun_out/r02_probe_hostcopy_n${N}.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n${N}.log 2> gpurun_out/r02_bench_n${N}.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02_bench_n${N}.err
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n${N}.log"):
    if l.startswith("{"):
        d=json.loads(l)
        print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "verify", d["verify"]["bit_exact_all_ranks"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_link_GBps_all_gpus"))
        for c in d["configs"]: print("  cfg", c["name"][:50], "ms/step", c["ms_per_step"], "value %.3e"%c["value"], c["parity"]["bit_exact"])
PY
