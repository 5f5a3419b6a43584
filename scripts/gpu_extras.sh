mkdir -p gpurun_out
timeout 500 python scripts/sweep_k1_extras.py > gpurun_out/sweep_k1_extras.log 2>&1; echo "rc=$?" >> gpurun_out/sweep_k1_extras.log
cat gpurun_out/sweep_k1_extras.log
