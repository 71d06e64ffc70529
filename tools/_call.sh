mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pool_fused.py tests/test_gpu_xfblock.py -m gpu -q -p no:cacheprovider > gpurun_out/r02af_pytest.log 2>&1
echo "pytest rc=$? : $(tail -1 gpurun_out/r02af_pytest.log)"; grep -E "^FAILED|^ERROR|Error|assert" gpurun_out/r02af_pytest.log | head -20
