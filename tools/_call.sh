timeout 900 python -m pytest tests/test_gpu_siglip.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --legs siglip_c2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_x.json; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); e=d['siglip_c2']; print(e.get('ms_per_step'), json.dumps(e.get('parity'))[:300])"
B200CLIP_GSTORE=0 timeout 600 python bench.py --steps 10 --warmup 3 --legs siglip_c2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_x.json; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); e=d['siglip_c2']; print(e.get('ms_per_step'))"
