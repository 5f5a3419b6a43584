mkdir -p gpurun_out
timeout 300 python scripts/r02_serial_ncu_target.py 65536 4 --noguard > gpurun_out/r02_serial_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_serial -c 1 -o gpurun_out/r02_c4_serial_v0_64k python scripts/r02_serial_ncu_target.py 65536 4 --noguard > gpurun_out/r02_serial_ncu.log 2>&1
tail -3 gpurun_out/r02_serial_ncu.log
