mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_clip_loss.py tests/test_gpu_kernels.py tests/test_gpu_siglip.py tests/test_gpu_gt_gemm.py tests/test_gpu_fullsize.py -q -p no:cacheprovider 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --legs none --no-cpu-baseline --no-graph 2>&1 | tail -1 > gpurun_out/bench_x.json; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print(d['ms_per_step'], d['parity']['loss_rel_vs_fp64'], d['parity']['grad_rel_sampled_rows'])"
