mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -x -q > gpurun_out/r02_pytest_serial.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_serial.log
tail -5 gpurun_out/r02_pytest_serial.log
timeout 900 python scripts/r02_c4_serial.py > gpurun_out/r02_c4_serial.jsonl 2>&1
cut -c1-420 gpurun_out/r02_c4_serial.jsonl
