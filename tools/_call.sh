mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clip_loss.py -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --legs none 2>&1 | tail -1 | tee gpurun_out/bench_gstore.json | cut -c1-1200
B200CLIP_GSTORE=0 timeout 600 python bench.py --steps 10 --warmup 3 --legs none 2>&1 | tail -1 | cut -c1-400
