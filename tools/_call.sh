mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_xfblock.py tests/test_gpu_tokens.py -x -q 2>&1 | tail -4
timeout 300 python tools/gpu_check_xfblock.py 2>&1 | grep -v Warning | tail -4
timeout 300 python tools/gpu_host_profile_agg.py 2>&1 | grep -v Warning > gpurun_out/agg_host.log; grep "host" gpurun_out/agg_host.log
