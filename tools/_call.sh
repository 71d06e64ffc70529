# scratch: the command of the last gpurun call of the round (full -m gpu suite + smoke on the final tree)
set -u
OUT=gpurun_out; TAG=r02e; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 $OUT/${TAG}_pytest_gpu.log)"
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$? : $(tail -1 $OUT/${TAG}_smoke.log)"
