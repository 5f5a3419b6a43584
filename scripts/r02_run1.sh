mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r02_pytest_gpu_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_2.log
tail -25 gpurun_out/r02_pytest_gpu_2.log
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_2.log 2> gpurun_out/r02_bench_2.err; echo "bench rc=$?" >> gpurun_out/r02_bench_2.err
tail -c 3000 gpurun_out/r02_bench_2.err
cut -c1-1500 gpurun_out/r02_bench_2.log
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r02_bench_ref_2.log 2> gpurun_out/r02_bench_ref_2.err; echo "ref rc=$?" >> gpurun_out/r02_bench_ref_2.err
tail -c 1000 gpurun_out/r02_bench_ref_2.err
cut -c1-1500 gpurun_out/r02_bench_ref_2.log
