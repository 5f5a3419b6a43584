mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log; tail -3 gpurun_out/r02_smoke.log
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r02_pytest_gpu_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_3.log
tail -14 gpurun_out/r02_pytest_gpu_3.log
# launch list of the bench command (lane count cut so that it fits under the profiler), incl. configs C1-C4 at reduced C4 steps
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-jit --lanes-per-gpu 131072 --c4-steps 200 > gpurun_out/r02_launches.log 2>&1; echo "launchlist rc=$?"
# K3: the stream kernel on C4, 2^20 lanes, 4 steps (2 rising)
timeout 300 python scripts/r02_stream_ncu_target.py 1048576 4 > gpurun_out/r02_stream_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_stream -c 1 -o gpurun_out/r02_k3_stream_1M python scripts/r02_stream_ncu_target.py 1048576 4 > gpurun_out/r02_stream_ncu.log 2>&1
tail -1 gpurun_out/r02_stream_ncu.log
# K4 / K5: signature, stimulus, pack / unpack, capture on a 20k-gate DAG
timeout 900 ncu --set full --clock-control none -k regex:"k_signature|k_stimulus|k_toggle|k_pack|k_unpack|k_capture|k_fill_rows" -c 12 -o gpurun_out/r02_k4_k5 python scripts/r02_k4k5_target.py > gpurun_out/r02_k4k5_ncu.log 2>&1
tail -1 gpurun_out/r02_k4k5_ncu.log
