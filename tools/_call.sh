mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clip_loss.py tests/test_gpu_kernels.py tests/test_gpu_z_alignment_diagnostics.py tests/test_gpu_siglip.py -m gpu -q -p no:cacheprovider > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02b_pytest.log)"
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider > gpurun_out/r02b_pytest_full.log 2>&1; echo "pytest full rc=$? $(tail -1 gpurun_out/r02b_pytest_full.log)"
timeout 300 python tools/gpu_probe_loss_error.py > gpurun_out/r02b_loss_error.json 2> gpurun_out/r02b_loss_error.err; echo "probe rc=$?"
timeout 600 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$? $(cut -c1-200 gpurun_out/r02b_bench.json)"
