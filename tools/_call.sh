mkdir -p gpurun_out
T=r02al
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 gpurun_out/${T}_pytest_gpu.log)"; grep -E "^FAILED|^ERROR" gpurun_out/${T}_pytest_gpu.log | head -20
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python tools/gpu_check_pool_tc.py 2>&1 | tail -3 | cut -c1-200
