mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "traces or lane_trick or c5 or c3 or level_kernel or streams or run_host" > gpurun_out/pytest_hints.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_hints.log
tail -3 gpurun_out/pytest_hints.log
timeout 900 python scripts/sweep.py --tiles 384,512,640 --streams 2 --hints 0,1 --unrolls 1,2,4 --caches 1 --blocks 128,256 --steps 3 > gpurun_out/sweep_hints.log 2>&1
cat gpurun_out/sweep_hints.log | cut -c1-250
