mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -x -q > gpurun_out/r02_pytest_serial.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_serial.log
tail -3 gpurun_out/r02_pytest_serial.log
timeout 900 python scripts/r02_c4_serial.py --quick > gpurun_out/r02_c4_serial_v2.jsonl 2>&1
cut -c1-330 gpurun_out/r02_c4_serial_v2.jsonl
timeout 300 python scripts/r02_serial_ncu_target.py 65536 2 > gpurun_out/r02_serial_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_serial -c 1 -o gpurun_out/r02_c4_serial_v2_64k python scripts/r02_serial_ncu_target.py 65536 2 > gpurun_out/r02_serial_ncu.log 2>&1
tail -2 gpurun_out/r02_serial_ncu.log
