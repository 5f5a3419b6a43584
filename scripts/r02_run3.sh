mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -x -q --durations=8 ) > gpurun_out/r02_pytest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_stream.log
tail -12 gpurun_out/r02_pytest_stream.log
timeout 900 python scripts/r02_c4_stream.py > gpurun_out/r02_c4_stream_b.jsonl 2> gpurun_out/r02_c4_stream.err
cut -c1-250 gpurun_out/r02_c4_stream_b.jsonl; tail -5 gpurun_out/r02_c4_stream.err
timeout 300 python scripts/r02_stream_ncu_target.py 1048576 4 > gpurun_out/r02_stream_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_stream -c 1 -o gpurun_out/r02_c4_stream_1M python scripts/r02_stream_ncu_target.py 1048576 4 > gpurun_out/r02_stream_ncu.log 2>&1
tail -2 gpurun_out/r02_stream_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_stream -c 1 -o gpurun_out/r02_c4_stream_64k python scripts/r02_stream_ncu_target.py 65536 4 > gpurun_out/r02_stream_ncu2.log 2>&1
tail -2 gpurun_out/r02_stream_ncu2.log
