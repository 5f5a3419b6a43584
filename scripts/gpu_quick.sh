mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline --verify > gpurun_out/bench_quick.log 2>&1; echo "rc=$?" >> gpurun_out/bench_quick.log
tail -c 900 gpurun_out/bench_quick.log
