mkdir -p gpurun_out
timeout 900 python scripts/sweep.py --tiles 256,384,512,640 --streams 2,3,4 --unrolls 2 --caches 1 --blocks 128 --steps 3 > gpurun_out/sweep_streams5.log 2>&1
cat gpurun_out/sweep_streams5.log | cut -c1-250
