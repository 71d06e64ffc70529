mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02e_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r02e_pytest_multi.log 2>&1; echo "pytest multi rc=$? $(tail -1 gpurun_out/r02e_pytest_multi.log)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err; echo "bench2 rc=$? $(cut -c1-250 gpurun_out/r02e_bench_n2.json)"; tail -3 gpurun_out/r02e_bench_n2.err
