#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md for this repo. Run under gpurun (1 GPU):
#   gpurun --timeout 900 -- 'bash tools/gpu_profile.sh r02 bw3_kernel'
# Produces in gpurun_out/: <tag>_bench_plain.json (plain run of the same command, must exit 0 first), <tag>_launches.csv
# (per-launch device times of the same command), <tag>_top.ncu-rep (--set full capture of the dominant kernel).
# Summaries are copied into profiles/ by tools/ncu_summarise.py.
set -u
TAG=${1:-r02}
KERNEL=${2:-bw3_kernel}
LEGS=${3:-none}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --legs $LEGS"
$CMD > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_launches.csv $CMD \
    > $OUT/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:${KERNEL} -s 6 -c 1 -f -o $OUT/${TAG}_top $CMD \
    > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la $OUT | tail -8
