mkdir -p gpurun_out
python tools/gpu_bench_tokens.py > gpurun_out/r02_tokens_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/r02_tokens_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:pool_fwd_mma -s 4 -c 1 -f -o gpurun_out/r02_poolfwd python tools/gpu_bench_tokens.py > gpurun_out/r02_poolfwd_ncu.log 2>&1; echo "fwd capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pool_bwd_mma -s 4 -c 1 -f -o gpurun_out/r02_poolbwd python tools/gpu_bench_tokens.py > gpurun_out/r02_poolbwd_ncu.log 2>&1; echo "bwd capture rc=$?"
