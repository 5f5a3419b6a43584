mkdir -p gpurun_out
timeout 100 python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > gpurun_out/bench_sanity.log 2>&1; echo "rc=$?" >> gpurun_out/bench_sanity.log
tail -c 600 gpurun_out/bench_sanity.log
