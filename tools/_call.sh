mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_retrieval.py -q -p no:cacheprovider 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --legs topk10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_x.json; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json'))
e=d['topk10']; print('topk10', e.get('value'), {kk:vv for kk,vv in e.items() if 'ms' in kk or 'parity' in kk})"
