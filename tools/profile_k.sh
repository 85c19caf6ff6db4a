#!/bin/bash
# `ncu --set full` capture of some kernels of the C2 bench + a launch list: bash tools/profile_k.sh <tag> <kernel-regex> [lines] [skip] [count]
TAG=${1:-k}
KRE=${2:-scan}
LINES=${3:-300000}
SKIP=${4:-3}
COUNT=${5:-1}
mkdir -p gpurun_out
CMD="python bench.py --lines $LINES --steps 1 --warmup 3 --e2e-lines 20000 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bvcf_ -s 60 -c 40 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $COUNT -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
