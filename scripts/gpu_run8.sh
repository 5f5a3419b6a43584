mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "persistent or traces or lane_trick" > gpurun_out/pytest_persist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_persist.log
tail -3 gpurun_out/pytest_persist.log
timeout 600 python scripts/sweep_persist.py > gpurun_out/sweep_persist.log 2>&1; echo "rc=$?" >> gpurun_out/sweep_persist.log
cat gpurun_out/sweep_persist.log
