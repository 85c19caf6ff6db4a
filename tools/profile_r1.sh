#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md: launch list + one full capture of the top kernels.
# usage (on the GPU box, via gpurun): bash tools/profile_r1.sh <tag> [lines]
TAG=${1:-r1}
LINES=${2:-300000}
mkdir -p gpurun_out
CMD="python bench.py --lines $LINES --steps 2 --warmup 3 --e2e-lines 20000 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bvcf_ -s 30 -c 15 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-900
tail -3 gpurun_out/ncu_f_$TAG.log
ls -la gpurun_out
