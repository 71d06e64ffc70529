mkdir -p gpurun_out
timeout 600 python bench.py --legs tokens_c3 --no-cpu-baseline > gpurun_out/r02ag_bench.json 2> gpurun_out/r02ag_bench.err; echo "bench rc=$?"
