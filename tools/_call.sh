mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02d_pytest.log)"
timeout 900 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$? $(cut -c1-300 gpurun_out/r02d_bench.json)"; tail -5 gpurun_out/r02d_bench.err
