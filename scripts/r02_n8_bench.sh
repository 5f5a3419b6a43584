mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n${N}_final.log 2> gpurun_out/r02_bench_n${N}_final.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02_bench_n${N}_final.err
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n${N}_final.log"):
    if l.startswith("{"):
        d=json.loads(l)
        print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "verify", d["verify"]["bit_exact_all_ranks"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_link_fraction_of_ceiling"))
        for c in d["configs"]: print("  cfg", c["name"][:50], "ms/step", c["ms_per_step"], "value %.3e"%c["value"], c["parity"]["bit_exact"])
PY
