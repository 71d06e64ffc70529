mkdir -p gpurun_out
T=r02m
timeout 300 python tools/gpu_check_xfblock.py > gpurun_out/${T}_xfblock.log 2>&1
echo "xfblock rc=$? : $(tail -14 gpurun_out/${T}_xfblock.log | cut -c1-400)"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 gpurun_out/${T}_pytest_gpu.log)"; grep -E "^FAILED|^ERROR" gpurun_out/${T}_pytest_gpu.log | head -20
