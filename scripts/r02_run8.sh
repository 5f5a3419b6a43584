mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "signature or toggle or stimulus or shards or smoke or capture or prevclock or checkpoint or traces" ) > gpurun_out/r02_pytest_k4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_k4.log
tail -5 gpurun_out/r02_pytest_k4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_signature|k_toggle" -c 8 --csv --log-file gpurun_out/r02_k4_after.csv python scripts/r02_k4k5_target.py > gpurun_out/r02_k4_after.log 2>&1
grep -E "k_toggle|k_signature" gpurun_out/r02_k4_after.csv | grep duration | cut -c1-260 | head -6
