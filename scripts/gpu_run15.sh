mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "resident or c4 or c1 or traces" > gpurun_out/pytest_res.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_res.log
tail -3 gpurun_out/pytest_res.log
timeout 600 python scripts/sweep_resident.py > gpurun_out/sweep_resident3.log 2>&1; echo "rc=$?" >> gpurun_out/sweep_resident3.log
cat gpurun_out/sweep_resident3.log | cut -c1-220
