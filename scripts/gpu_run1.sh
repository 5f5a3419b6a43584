mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --gates 100000 --lanes-per-gpu 1048576 --steps 3 --warmup 3 --verify --cpu-seconds 3 > gpurun_out/bench_small.log 2>&1; echo "rc=$?" >> gpurun_out/bench_small.log
timeout 600 python bench.py --steps 3 --warmup 3 --verify > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
tail -3 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench_small.log gpurun_out/bench_full.log
