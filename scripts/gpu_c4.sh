mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "c4" > gpurun_out/pytest_c4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_c4.log
tail -n 3 gpurun_out/pytest_c4.log
timeout 300 python scripts/c4_share.py > gpurun_out/c4_share.log 2>&1; cat gpurun_out/c4_share.log
