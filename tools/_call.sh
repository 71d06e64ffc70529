mkdir -p gpurun_out
T=r02u
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 gpurun_out/${T}_pytest_gpu.log)"; grep -E "^FAILED|^ERROR" gpurun_out/${T}_pytest_gpu.log | head -20
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err ) 2> gpurun_out/${T}_bench.time
echo "bench rc=$? $(tail -3 gpurun_out/${T}_bench.time | head -1)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$? $(cut -c1-200 gpurun_out/${T}_bench_ref.json)"
