mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python scripts/sweep_resident.py > gpurun_out/sweep_resident2.log 2>&1; echo "rc=$?" >> gpurun_out/sweep_resident2.log
cat gpurun_out/sweep_resident2.log | cut -c1-200
timeout 600 python scripts/bench_configs.py c1 c2 c4 > gpurun_out/bench_configs3.log 2>&1; echo "rc=$?" >> gpurun_out/bench_configs3.log
grep -v cpu_port gpurun_out/bench_configs3.log | cut -c1-330
