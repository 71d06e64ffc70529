mkdir -p gpurun_out
T=r02ak
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29542 tools/gpu_check_dist.py > gpurun_out/${T}_dist_check_n${NG}.log 2>&1
echo "dist check n$NG rc=$? : $(tail -1 gpurun_out/${T}_dist_check_n${NG}.log | cut -c1-300)"; grep -i "mismatch\|error\|timeout" gpurun_out/${T}_dist_check_n${NG}.log | head -8
timeout 300 $TR --master-port 29541 bench.py --gpus $NG --no-cpu-baseline > gpurun_out/${T}_bench_n${NG}.json 2> gpurun_out/${T}_bench_n${NG}.err
echo "bench n$NG rc=$? : $(cut -c1-120 gpurun_out/${T}_bench_n${NG}.json)"; tail -2 gpurun_out/${T}_bench_n${NG}.err | cut -c1-200
