mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --verify > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log
tail -c 600 gpurun_out/bench_full.log; tail -c 400 gpurun_out/bench_ref.log
