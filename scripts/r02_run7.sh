mkdir -p gpurun_out
timeout 900 python scripts/r02_c4_stream.py --quick > gpurun_out/r02_c4_stream_f.jsonl 2> gpurun_out/r02_c4_stream.err
cut -c1-250 gpurun_out/r02_c4_stream_f.jsonl; tail -5 gpurun_out/r02_c4_stream.err
( time timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -x -q -k "many_steps or c4_shape or guard or tiled" ) > gpurun_out/r02_pytest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_stream.log
tail -4 gpurun_out/r02_pytest_stream.log
timeout 300 python scripts/r02_stream_ncu_target.py 65536 4 > gpurun_out/r02_stream_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_stream -c 1 -o gpurun_out/r02_c4_stream32_64k python scripts/r02_stream_ncu_target.py 65536 4 > gpurun_out/r02_stream_ncu2.log 2>&1
tail -2 gpurun_out/r02_stream_ncu2.log
