set -u
OUT=gpurun_out; TAG=r02b; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 $OUT/${TAG}_pytest_gpu.log)"
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$? : $(tail -1 $OUT/${TAG}_smoke.log)"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$? : $(cut -c1-200 $OUT/${TAG}_bench.json)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
echo "ref rc=$? : $(cut -c1-300 $OUT/${TAG}_bench_ref.json)"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --legs none"
$CMD > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"(bw3_kernel|gt_gemm_kernel)" -s 4 -c 2 -f -o $OUT/${TAG}_top $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
