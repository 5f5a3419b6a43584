mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
free -g | head -2 > gpurun_out/host_n${N}.txt; nproc >> gpurun_out/host_n${N}.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n${N}.log
tail -c 900 gpurun_out/bench_n${N}.log
