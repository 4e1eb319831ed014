#!/bin/bash
# ncu passes of B200_PROFILING.md: launch list of one bench step, then --set full on the top kernels
# usage: gpurun_ncu.sh <tag> [kernel-regex]
set -x
TAG=${1:-r1}
PAT=${2:-"gat_aggregate_tile|proj_umma"}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --quick"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 68 -c 17 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s 24 -c 8 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
ls -la gpurun_out
