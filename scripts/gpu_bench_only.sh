mkdir -p gpurun_out
timeout 600 python bench.py --verify > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
tail -c 700 gpurun_out/bench_full.log
