mkdir -p gpurun_out
timeout 300 python scripts/r02_stream_ncu_target.py 1048576 4 > gpurun_out/r02_stream_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_stream -c 1 -o gpurun_out/r02_k3_stream_1M_final python scripts/r02_stream_ncu_target.py 1048576 4 > gpurun_out/r02_stream_ncu.log 2>&1
tail -1 gpurun_out/r02_stream_ncu.log
