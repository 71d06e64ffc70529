mkdir -p gpurun_out
T=r02f
timeout 300 python tools/gpu_check_pool_tc.py > gpurun_out/${T}_pool_tc.log 2>&1
echo "pool tc rc=$? : $(tail -10 gpurun_out/${T}_pool_tc.log | cut -c1-250)"
timeout 300 python tools/gpu_check_pool_tc.py --fp16 > gpurun_out/${T}_pool_tc_fp16.log 2>&1
echo "pool tc fp16 rc=$? : $(tail -3 gpurun_out/${T}_pool_tc_fp16.log | cut -c1-250)"
timeout 300 python tools/gpu_bench_pool_tc.py 32 3136 512 4 9 > gpurun_out/${T}_pool_tc_bench.log 2>&1
echo "bench rc=$? : $(tail -3 gpurun_out/${T}_pool_tc_bench.log | cut -c1-330)"
