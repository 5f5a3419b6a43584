mkdir -p gpurun_out
timeout 600 python scripts/r02_c4_stream.py --quick > gpurun_out/r02_c4_stream_h.jsonl 2> gpurun_out/r02_c4_stream.err
cut -c1-250 gpurun_out/r02_c4_stream_h.jsonl | head -2
timeout 800 python scripts/r02_modes_compare.py > gpurun_out/r02_modes_compare_c.jsonl 2> gpurun_out/r02_modes_compare.err; python -c "
import json
for l in open('gpurun_out/r02_modes_compare_c.jsonl'):
    d=json.loads(l); print(d['design'][:30], d['lanes'], 'levels', d['levels'], 'resident', d['resident'], 'serial', d['serial'], 'auto', d['auto'], d['auto_mode'])
"
( time timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -x -q ) > gpurun_out/r02_pytest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_stream.log
tail -4 gpurun_out/r02_pytest_stream.log
