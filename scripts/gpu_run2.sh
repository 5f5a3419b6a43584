mkdir -p gpurun_out
timeout 900 python scripts/sweep.py --tiles 256,512,1024,2048,4096,8192,16384 --unrolls 4 --caches 1 > gpurun_out/sweep_tiles.log 2>&1
timeout 600 python scripts/sweep.py --tiles 1024,4096 --unrolls 1,2,4,8 --caches 0,1 --steps 1 > gpurun_out/sweep_variants.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --lanes-per-gpu 1048576"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_eval_level -s 300 -c 3 -o gpurun_out/prof_eval_level $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
