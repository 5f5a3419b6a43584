mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 2 gpurun_out/smoke.log; tail -n 3 gpurun_out/pytest_gpu.log
CMD2="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --lanes-per-gpu 32768"
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_eval_level -s 200 -c 200 --csv --log-file gpurun_out/ncu_steady_T512.csv $CMD2 > gpurun_out/ncu_steady.log 2>&1
ls -la gpurun_out/ncu_steady_T512.csv
