mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
tail -c 1300 gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log
CMD2="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --lanes-per-gpu 32768"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval_level -s 250 -c 3 -o gpurun_out/prof_eval_level_default $CMD2 > gpurun_out/ncu_full_default.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --lanes-per-gpu 131072"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_launches_default.log 2>&1
ls -la gpurun_out | grep -E "launches_default|prof_eval_level_default"
