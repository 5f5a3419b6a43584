mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -x -q -k "many_steps or c4_shape or guard or tiled or traces" ) > gpurun_out/r02_pytest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_stream.log
tail -6 gpurun_out/r02_pytest_stream.log
timeout 900 python scripts/r02_c4_stream.py > gpurun_out/r02_c4_stream_c.jsonl 2> gpurun_out/r02_c4_stream.err
cut -c1-250 gpurun_out/r02_c4_stream_c.jsonl; tail -5 gpurun_out/r02_c4_stream.err
timeout 300 python scripts/r02_stream_ncu_target.py 1048576 4 > gpurun_out/r02_stream_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_stream -c 1 -o gpurun_out/r02_c4_stream_1M_b python scripts/r02_stream_ncu_target.py 1048576 4 > gpurun_out/r02_stream_ncu.log 2>&1
tail -2 gpurun_out/r02_stream_ncu.log
