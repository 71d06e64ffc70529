mkdir -p gpurun_out
T=r02k
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 bench.py --gpus 8 > gpurun_out/${T}_bench_n8.json 2> gpurun_out/${T}_bench_n8.err
echo "bench n8 rc=$? : $(cut -c1-300 gpurun_out/${T}_bench_n8.json)"
timeout 300 $TR --master-port 29542 tools/gpu_check_dist.py > gpurun_out/${T}_dist_check_n8.log 2>&1
echo "dist check n8 rc=$? : $(tail -2 gpurun_out/${T}_dist_check_n8.log | cut -c1-300)"
B200CLIP_SYMM=0 timeout 300 $TR --master-port 29543 bench.py --gpus 8 --legs none --no-cpu-baseline > gpurun_out/${T}_bench_n8_nccl.json 2> gpurun_out/${T}_bench_n8_nccl.err
echo "bench n8 nccl rc=$? : $(cut -c1-300 gpurun_out/${T}_bench_n8_nccl.json)"
