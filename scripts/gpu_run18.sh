mkdir -p gpurun_out
timeout 900 python scripts/sweep.py --tiles 768,1024,1280,1536 --streams 3,4 --unrolls 2 --caches 1,0 --blocks 128 > gpurun_out/sweep_streams2.log 2>&1
cat gpurun_out/sweep_streams2.log | cut -c1-250
