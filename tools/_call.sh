mkdir -p gpurun_out
T=r02i
timeout 300 python tools/gpu_check_pool_fused.py > gpurun_out/${T}_pool_fused.log 2>&1
echo "pool fused rc=$? : $(tail -12 gpurun_out/${T}_pool_fused.log | cut -c1-400)"
timeout 300 python tools/gpu_check_pool_tc.py > gpurun_out/${T}_pool_tc.log 2>&1
echo "pool tc rc=$? : $(tail -2 gpurun_out/${T}_pool_tc.log | cut -c1-250)"
