mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resident_word" > gpurun_out/pytest_res.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_res.log
tail -3 gpurun_out/pytest_res.log
timeout 900 python scripts/sweep_resident.py > gpurun_out/sweep_resident.log 2>&1; echo "rc=$?" >> gpurun_out/sweep_resident.log
cat gpurun_out/sweep_resident.log | cut -c1-200
