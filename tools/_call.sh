mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_z_topk_two_sweeps.py tests/test_gpu_kernels.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --legs retrieval,topk10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_x.json; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json'))
for k in ('retrieval','topk10'):
    e=d[k]; print(k, e.get('value'), {kk:vv for kk,vv in e.items() if 'ms' in kk or 'exact' in kk or 'parity' in kk})"
