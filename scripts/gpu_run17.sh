mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "several_streams or run_host" > gpurun_out/pytest_streams.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_streams.log
tail -3 gpurun_out/pytest_streams.log
timeout 900 python scripts/sweep.py --tiles 2048,1024,512 --streams 1,2,3 --unrolls 2 --caches 1 --blocks 128 > gpurun_out/sweep_streams.log 2>&1
cat gpurun_out/sweep_streams.log | cut -c1-250
