mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_gt_gemm.py tests/test_gpu_milpool.py -x -q -p no:cacheprovider > gpurun_out/r02c_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/r02c_memcheck.log | head -10
