mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum"
for T in 256 512; do for H in 1; do
  timeout 400 ncu --cache-control none --clock-control none --metrics $M -k regex:k_eval_persistent -s 0 -c 3 --csv --log-file gpurun_out/ncu_persist_${T}_${H}.csv python scripts/persist_one.py $T $H > gpurun_out/ncu_p_${T}_${H}.log 2>&1
done; done
cat gpurun_out/ncu_persist_256_1.csv gpurun_out/ncu_persist_512_1.csv | cut -c1-400
