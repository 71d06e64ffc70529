mkdir -p gpurun_out
T=r02v
timeout 600 python -m pytest tests/test_gpu_siglip.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$? : $(tail -1 gpurun_out/${T}_pytest.log)"; grep -E "^FAILED|^ERROR" gpurun_out/${T}_pytest.log | head
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_siglip_launches.csv python tools/gpu_siglip_step.py 4 > gpurun_out/${T}_ncu1.log 2>&1; echo "siglip list rc=$?"
grep -E "siglip_compact|siglip_pos" gpurun_out/${T}_siglip_launches.csv | tail -2 | cut -c1-70,200-400
