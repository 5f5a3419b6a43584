mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --lanes-per-gpu 131072"
$CMD > gpurun_out/plain_small.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_level -s 250 -c 3 -o gpurun_out/prof_eval_level_default $CMD > gpurun_out/ncu_full_default.log 2>&1
$CMD > gpurun_out/plain_small2.log 2>&1 && timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_eval_level -s 200 -c 200 --csv --log-file gpurun_out/ncu_nocachectl.csv $CMD > gpurun_out/ncu_nocachectl.log 2>&1
tail -2 gpurun_out/pytest_gpu.log; ls -la gpurun_out | grep -E "prof_eval_level_default|nocachectl"
