mkdir -p gpurun_out
T=r02q
timeout 300 python tools/gpu_check_xfblock.py > gpurun_out/${T}_xfblock.log 2>&1
echo "xfblock rc=$? : $(tail -4 gpurun_out/${T}_xfblock.log | cut -c1-300)"
timeout 300 python tools/gpu_check_pool_fused.py > gpurun_out/${T}_pool_fused.log 2>&1
echo "pool fused rc=$? : $(tail -4 gpurun_out/${T}_pool_fused.log | cut -c1-300)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_tokens_launches.csv python tools/gpu_tokens_step.py 4 > gpurun_out/${T}_ncu.log 2>&1
echo "launch list rc=$?"
