mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gt_gemm.py tests/test_gpu_clip_loss.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --legs clip32k_d768 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_gstore.json; python -c "
import json; d=json.load(open('gpurun_out/bench_gstore.json')); print(d['ms_per_step']); e=d['clip32k_d768']; print(e.get('ms_per_step'), json.dumps(e.get('parity'))[:300], json.dumps(e.get('roofline',{}).get('backward_one_recompute'))[150:600])"
