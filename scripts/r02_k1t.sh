mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "level_tma" ) > gpurun_out/r02_pytest_k1t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_k1t.log
tail -4 gpurun_out/r02_pytest_k1t.log
for t in "level_tma=1" "level_tma=1,level_pdl=0"; do
  timeout 600 python bench.py --steps 3 --warmup 2 --no-configs --no-jit --no-e2e --no-cpu-baseline --tune $t > gpurun_out/r02_k1t2_bench_$t.json 2> gpurun_out/r02_k1t_bench.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_k1t2_bench_$t.json').read().strip().splitlines()[-1])
print('$t', 'value %.4e'%d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['verify']['bit_exact'], d['clocks'])
" || tail -5 gpurun_out/r02_k1t_bench.err
done
