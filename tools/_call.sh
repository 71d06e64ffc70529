mkdir -p gpurun_out
T=r02d
timeout 300 python tools/gpu_check_pool_tc.py > gpurun_out/${T}_pool_tc.log 2>&1
echo "pool tc rc=$? : $(tail -8 gpurun_out/${T}_pool_tc.log | cut -c1-250)"
timeout 300 python tools/gpu_check_pool_tc.py --fp16 > gpurun_out/${T}_pool_tc_fp16.log 2>&1
echo "pool tc fp16 rc=$? : $(tail -4 gpurun_out/${T}_pool_tc_fp16.log | cut -c1-250)"
