mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_milpool.py -x -q 2>&1 | tail -8
timeout 600 python tools/gpu_check_milpool.py 2>&1 | grep -v Warning | tee gpurun_out/r02_milpool_check.log | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mil_(gate_fwd|dx|dw|dpre|pool_fwd|dA)_kernel" -c 6 -f -o gpurun_out/r02_milpool python tools/gpu_milpool_step.py 128 1 > gpurun_out/ncu_mil.log 2>&1; tail -3 gpurun_out/ncu_mil.log
