mkdir -p gpurun_out
timeout 600 python scripts/r02_c4_stream.py --quick > gpurun_out/r02_c4_stream_l.jsonl 2> gpurun_out/r02_c4_stream.err
cut -c1-250 gpurun_out/r02_c4_stream_l.jsonl | head -2; tail -3 gpurun_out/r02_c4_stream.err
( time timeout 900 python -m pytest tests/test_gpu_serial.py tests/test_gpu_fullsize.py -m gpu -x -q -k "not c5_one and not c3_mult" ) > gpurun_out/r02_pytest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_stream.log
tail -4 gpurun_out/r02_pytest_stream.log
