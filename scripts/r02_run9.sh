mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r02_pytest_gpu_5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_4.log
tail -12 gpurun_out/r02_pytest_gpu_5.log
timeout 600 python scripts/r02_c4_stream.py  > gpurun_out/r02_c4_stream_k.jsonl 2> gpurun_out/r02_c4_stream.err
cut -c1-250 gpurun_out/r02_c4_stream_k.jsonl | head -3
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_4.log 2> gpurun_out/r02_bench_4.err; echo "bench rc=$?" >> gpurun_out/r02_bench_4.err
tail -c 300 gpurun_out/r02_bench_4.err
