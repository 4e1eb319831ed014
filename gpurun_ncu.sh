#!/bin/bash
# ncu passes of B200_PROFILING.md: launch list of one bench step, then --set full on the top kernels
# usage: gpurun_ncu.sh <tag> [kernel-regex]
set -x
TAG=${1:-r1}
PAT=${2:-"gat_aggregate_tile|proj_umma|proj_pair|gb_emit"}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --quick"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s ${SKIP:-30} -c ${COUNT:-10} -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
