mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -x -q --durations=8 ) > gpurun_out/r02_pytest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_stream.log
tail -30 gpurun_out/r02_pytest_stream.log
timeout 900 python scripts/r02_c4_stream.py > gpurun_out/r02_c4_stream.jsonl 2> gpurun_out/r02_c4_stream.err
cut -c1-300 gpurun_out/r02_c4_stream.jsonl; tail -5 gpurun_out/r02_c4_stream.err
