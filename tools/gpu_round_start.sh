#!/bin/bash
# First GPU call of a round: everything that was written without hardware gets its verdict in ONE box session.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_start.sh r02'
# Writes into gpurun_out/: <tag>_pytest_gpu.log (the whole -m gpu suite, no -x so every failure shows), <tag>_smoke.log,
# <tag>_bench.json (+ .err), <tag>_multipos.log, <tag>_clspool.log, <tag>_tokens.log. Nothing here runs under a profiler.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
B200CLIP_RUN_UNVERIFIED=1 timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$? : $(tail -1 $OUT/${TAG}_pytest_gpu.log)"
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$? : $(tail -1 $OUT/${TAG}_smoke.log)"
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$? : $(cut -c1-300 $OUT/${TAG}_bench.json)"
B200CLIP_POOL_FUSED_DQ=1 timeout 300 python tools/gpu_bench_tokens.py > $OUT/${TAG}_tokens_fused_dq.log 2>&1
echo "tokens (fused dq) rc=$? : $(tail -2 $OUT/${TAG}_tokens_fused_dq.log | cut -c1-300)"
for t in multipos clspool tokens topk; do
  timeout 300 python tools/gpu_bench_${t}.py > $OUT/${TAG}_${t}.log 2>&1
  echo "$t rc=$? : $(tail -2 $OUT/${TAG}_${t}.log | cut -c1-300)"
done
