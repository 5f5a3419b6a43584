mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_serial.py -m gpu -x -q -k "c4_shape or guard or tiled or traces" > gpurun_out/r02_pytest_serial.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_serial.log
tail -3 gpurun_out/r02_pytest_serial.log
timeout 900 python scripts/r02_c4_serial_pf.py > gpurun_out/r02_c4_serial_pf.jsonl 2>&1
cat gpurun_out/r02_c4_serial_pf.jsonl
