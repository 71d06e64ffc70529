mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02c_pytest.log)"
timeout 300 python tools/gpu_probe_precision.py > gpurun_out/r02c_precision_probe.json 2> gpurun_out/r02c_precision_probe.err; echo "probe rc=$?"
timeout 600 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$? $(cut -c1-200 gpurun_out/r02c_bench.json)"
