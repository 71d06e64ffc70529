set -u
OUT=gpurun_out; TAG=r02c; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 tools/gpu_check_dist.py > $OUT/${TAG}_dist4.log 2>&1
echo "dist check rc=$? : $(tail -1 $OUT/${TAG}_dist4.log | cut -c1-200)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n4.json 2> $OUT/${TAG}_bench_n4.err
echo "bench4 rc=$? : $(cut -c1-250 $OUT/${TAG}_bench_n4.json)"
