# scratch: the command of the last gpurun call of the round (full -m gpu suite, smoke, bench on the final tree)
set -u
OUT=gpurun_out; TAG=r02d; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 $OUT/${TAG}_pytest_gpu.log)"
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$? : $(tail -1 $OUT/${TAG}_smoke.log)"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$? : $(cut -c1-200 $OUT/${TAG}_bench.json)"
