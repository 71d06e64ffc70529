mkdir -p gpurun_out
T=r02ae
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 600 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench n1 rc=$?"
timeout 300 $TR --master-port 29541 bench.py --gpus $NG --no-cpu-baseline > gpurun_out/${T}_bench_n${NG}.json 2> gpurun_out/${T}_bench_n${NG}.err
echo "bench n$NG rc=$? : $(cut -c1-200 gpurun_out/${T}_bench_n${NG}.json)"
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 gpurun_out/${T}_pytest_gpu.log)"; grep -E "^FAILED|^ERROR" gpurun_out/${T}_pytest_gpu.log | head
