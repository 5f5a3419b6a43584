mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python scripts/sweep.py --tiles 1024,2048,4096 --unrolls 2 --caches 0,1 --blocks 128,256 > gpurun_out/sweep_u2.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --verify > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --lanes-per-gpu 262144 --tile-words 1024"
$CMD > gpurun_out/plain_t1024.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:k_eval_level -s 1000 -c 250 --csv --log-file gpurun_out/ncu_t1024.csv $CMD > gpurun_out/ncu_t1024.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
