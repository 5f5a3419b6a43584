mkdir -p gpurun_out
timeout 900 python scripts/r02_c4_stream.py  > gpurun_out/r02_c4_stream_e.jsonl 2> gpurun_out/r02_c4_stream.err
cut -c1-250 gpurun_out/r02_c4_stream_e.jsonl; tail -5 gpurun_out/r02_c4_stream.err
( time timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -x -q -k "many_steps or c4_shape or guard or tiled" ) > gpurun_out/r02_pytest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_stream.log
tail -4 gpurun_out/r02_pytest_stream.log
