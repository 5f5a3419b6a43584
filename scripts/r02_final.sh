mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log; tail -2 gpurun_out/r02_smoke.log
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_final.log
tail -12 gpurun_out/r02_pytest_gpu_final.log
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r02_bench_ref_final.log 2> gpurun_out/r02_bench_ref_final.err; echo "ref rc=$?"
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_final.log 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?" >> gpurun_out/r02_bench_final.err
tail -c 300 gpurun_out/r02_bench_final.err
