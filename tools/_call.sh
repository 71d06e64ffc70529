mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_milpool.py -q 2>&1 | tail -5
timeout 600 python tools/gpu_check_milpool.py 2>&1 | grep -v Warning | tee gpurun_out/r02_milpool_check.log | tail -10
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"(te_kernel|gt_gemm_kernel|mil_split_x|mil_dpre_tc)" -c 7 -f -o gpurun_out/r02_milpool_tc python tools/gpu_milpool_step.py 128 1 > gpurun_out/ncu_mil.log 2>&1; tail -2 gpurun_out/ncu_mil.log
