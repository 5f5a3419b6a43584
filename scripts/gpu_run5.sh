mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 12700 -c 13000 --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_launches_default.log 2>&1
$CMD > gpurun_out/plain_default2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_eval_level -s 15000 -c 3 -o gpurun_out/prof_eval_level_default $CMD > gpurun_out/ncu_full_default.log 2>&1
$CMD > gpurun_out/plain_default3.log 2>&1 && ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_eval_level -s 15000 -c 400 --csv --log-file gpurun_out/ncu_nocachectl.csv $CMD > gpurun_out/ncu_nocachectl.log 2>&1
ls -la gpurun_out | tail -12
