mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/r02_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_0.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_0.log
timeout 600 python scripts/r02_probe_l2_ceiling.py > gpurun_out/r02_probe_l2_ceiling.jsonl 2>&1
timeout 300 python scripts/r02_probe_hostcopy.py > gpurun_out/r02_probe_hostcopy_n1.jsonl 2>&1
timeout 300 python scripts/r02_c4_ncu_target.py > gpurun_out/r02_c4_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_resident -c 1 -o gpurun_out/r02_c4_resident_before python scripts/r02_c4_ncu_target.py > gpurun_out/r02_c4_ncu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu_0.log; cat gpurun_out/r02_probe_l2_ceiling.jsonl gpurun_out/r02_probe_hostcopy_n1.jsonl; tail -3 gpurun_out/r02_c4_ncu.log
