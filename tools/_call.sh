mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gt_gemm.py tests/test_gpu_clip_loss.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --legs none 2>&1 | tail -1 > gpurun_out/bench_gstore.json; python -c "
import json; d=json.load(open('gpurun_out/bench_gstore.json')); print(d['ms_per_step'], d['roofline']['ms_per_launch'], d['parity']['grad_rel_sampled_rows'], json.dumps(d['roofline'].get('backward_one_recompute'))[200:560])"
