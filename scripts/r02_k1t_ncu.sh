mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval_level_tma -s 250 -c 2 -o gpurun_out/r02_k1t python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-jit --no-verify --lanes-per-gpu 32768 --tune level_tma=1,level_pdl=0 > gpurun_out/r02_k1t_ncu.log 2>&1
tail -2 gpurun_out/r02_k1t_ncu.log
