mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "persistent" > gpurun_out/pytest_persist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_persist.log
tail -2 gpurun_out/pytest_persist.log
timeout 900 python scripts/sweep_persist.py --tiles 256,512,1024 --unrolls 2 --hints 0,1 --ctas 1,2 --blocks 512,1024 > gpurun_out/sweep_persist2.log 2>&1; echo "rc=$?" >> gpurun_out/sweep_persist2.log
cat gpurun_out/sweep_persist2.log
