mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "persistent or traces or lane_trick" > gpurun_out/pytest_persist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_persist.log
tail -3 gpurun_out/pytest_persist.log
timeout 900 python scripts/sweep_persist.py --tiles 256,512,1024 --inflight 1,2,3 --unrolls 2 --hints 1 --ctas 2 --blocks 512 > gpurun_out/sweep_pipe.log 2>&1; echo "rc=$?" >> gpurun_out/sweep_pipe.log
cat gpurun_out/sweep_pipe.log | cut -c1-260
