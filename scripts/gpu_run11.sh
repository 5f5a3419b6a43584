mkdir -p gpurun_out
timeout 900 python scripts/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "rc=$?" >> gpurun_out/bench_configs.log
cat gpurun_out/bench_configs.log
